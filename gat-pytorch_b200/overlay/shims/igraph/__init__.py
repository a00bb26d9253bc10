"""Import-only stand-in for python-igraph (absent from the offline image), used by
visualisation/neighbourhood_attention_weights.py: builds the attribute bags the script fills, draws nothing."""


class _Seq(dict):
    pass


class Graph:
    def __init__(self, *args, **kwargs):
        self.vs, self.es, self._n, self._edges = _Seq(), _Seq(), 0, []

    def add_vertices(self, n):
        self._n += int(n)

    def add_edges(self, edges):
        self._edges.extend(list(edges))

    def __getattr__(self, name):          # layout_reingold_tilford_circular(...) and friends: accepted, nothing computed
        if name.startswith("layout") or name.startswith("delete") or name.startswith("simplify"):
            return lambda *args, **kwargs: None
        raise AttributeError(name)


class _Plot:
    def save(self, *args, **kwargs):
        pass


def plot(*args, **kwargs):
    return _Plot()
