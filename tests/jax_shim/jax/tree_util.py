"""`jax.tree_util.tree_map` over tuples / lists / dicts / NamedTuples."""


def tree_map(f, tree, *rest):
    from . import _tree_map
    return _tree_map(f, tree, *rest)


def tree_leaves(tree):
    from . import _leaves
    return _leaves(tree)
