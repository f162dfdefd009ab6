"""grakel.graph.Graph stand-in: holds a directed edge dictionary and node labels.

Semantics relied on by the reference (grakel_replace/weisfeiler_lehman.py:193-214):
`get_edge_dictionary()` returns {v: {successor: weight}} with a key for every vertex that
is an endpoint of some edge; `get_labels(purpose="dictionary")` returns {node: label}.
"""


class Graph:
    def __init__(self, initialization_object=None, node_labels=None, edge_labels=None,
                 graph_format="auto", construct_labels=False):
        ed = {}
        obj = initialization_object or {}
        for key, val in obj.items():
            if isinstance(key, tuple):
                u, v = key
                ed.setdefault(u, {})[v] = val
                ed.setdefault(v, {})
            else:
                ed.setdefault(key, {})
                for v, w in dict(val).items():
                    ed[key][v] = w
                    ed.setdefault(v, {})
        self._ed = ed
        self._nl = dict(node_labels) if node_labels is not None else {}
        self._el = edge_labels

    def desired_format(self, *_a, **_k):
        return None

    def get_edge_dictionary(self):
        return self._ed

    def get_labels(self, purpose="dictionary", label_type="vertex", return_none=False):
        if label_type == "edge":
            if not self._el:
                return None if return_none else {}
            return self._el
        return self._nl
