"""Pick the strongest oracle available: the unmodified reference hot path
(oracle/_ref/libmdzref.so) when it is present, else the C restatement
(oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY."""
import refpath

_ref = None


def oracle_render(view, threads=None):
    global _ref
    if _ref is None:
        _ref = refpath.load() or False
    if _ref:
        raw, _ = refpath.ref_render(_ref, view, threads)
        return raw, "reference"
    import portpath
    return portpath.port_render(view), "port"
