class Semiring(object):
    """Base class the reference's GraphSemiring derives from (utils/graph_semiring.py:2)."""
