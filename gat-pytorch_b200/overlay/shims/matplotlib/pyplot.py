"""Records the calls (`CALLS`) instead of drawing; `gcf().savefig(path)` notes the path it would have written."""
CALLS = []


class _Figure:
    def savefig(self, path, *args, **kwargs):
        CALLS.append(("savefig", str(path)))


def _recorder(name):
    def fn(*args, **kwargs):
        CALLS.append((name, len(args)))
    return fn


bar, hist, xlabel, ylabel, title, legend, show, close, figure, plot, yscale, xscale, tight_layout = (
    _recorder(n) for n in ("bar", "hist", "xlabel", "ylabel", "title", "legend", "show", "close", "figure", "plot", "yscale", "xscale",
                           "tight_layout"))


def gcf():
    return _Figure()
