/* oracle/shim/glib.h -- the few glib scalar typedefs the non-GUI reference
 * sources use, so they compile headless (no GTK2 on this box).
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_GLIB_H
#define ORACLE_SHIM_GLIB_H
#include <stdint.h>
#include <limits.h>
#include <stdlib.h>
#include <stdbool.h>
typedef int            gboolean;
typedef int            gint;
typedef unsigned int   guint;
typedef uint32_t       guint32;
typedef uint16_t       guint16;
typedef uint8_t        guint8;
typedef unsigned char  guchar;
typedef char           gchar;
typedef double         gdouble;
typedef float          gfloat;
typedef void*          gpointer;
typedef long           glong;
typedef unsigned long  gulong;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#define G_LITTLE_ENDIAN 1234
#define G_BIG_ENDIAN    4321
#define G_BYTE_ORDER    G_LITTLE_ENDIAN
#define g_free free
/* opaque GTK/GDK handles: only ever used as pointers in headers we pass through */
typedef struct _ShimGtkWidget GtkWidget;
typedef struct _ShimGtkWindow GtkWindow;
typedef struct _ShimGtkBox GtkBox;
typedef struct _ShimGtkButton GtkButton;
typedef struct _ShimGtkObject GtkObject;
typedef struct _ShimGtkAdjustment GtkAdjustment;
typedef struct _ShimGtkToggleButton GtkToggleButton;
typedef struct _ShimGtkEditable GtkEditable;
typedef struct _ShimGtkSpinButton GtkSpinButton;
typedef struct _ShimGdkEventButton GdkEventButton;
typedef struct _ShimGdkEventMotion GdkEventMotion;
typedef struct _ShimGdkEventExpose GdkEventExpose;
typedef struct _ShimGdkEventConfigure GdkEventConfigure;
typedef struct _ShimGdkEventCrossing GdkEventCrossing;
typedef struct _ShimGdkEvent GdkEvent;
typedef struct _ShimGdkGC GdkGC;
typedef struct _ShimGdkPixmap GdkPixmap;
typedef struct _ShimGdkColor GdkColor;
#endif
