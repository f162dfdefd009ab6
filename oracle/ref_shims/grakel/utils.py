"""grakel.utils.graph_from_networkx stand-in (call sites kernels/weisfilerlehman.py:134-136)."""


def graph_from_networkx(X, node_labels_tag=None, edge_labels_tag=None, edge_weight_tag=None,
                        as_Graph=False):
    for g in X:
        if not g.is_directed():
            g = g.to_directed()
        edges = {(u, v): 1 for (u, v) in g.edges()}
        nl = {n: d.get(node_labels_tag) for n, d in g.nodes(data=True)} \
            if node_labels_tag is not None else None
        el = None
        if edge_labels_tag is not None:
            el = {(u, v): d.get(edge_labels_tag) for u, v, d in g.edges(data=True)}
        yield [edges, nl, el]
