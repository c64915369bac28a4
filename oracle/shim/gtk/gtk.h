/* oracle/shim/gtk/gtk.h -- see ../glib.h.  TEST INFRASTRUCTURE ONLY. */
#include "../glib.h"
