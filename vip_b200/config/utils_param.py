"""kwargs <-> parameter-object plumbing (``vip_hci/config/utils_param.py:60-164``)."""
from inspect import signature

# kwargs that always belong to the parameter object even if it has no such field
_ALWAYS_CLASS_KEYS = ("param",)


def separate_kwargs_dict(initial_kwargs, parent_class):
    """Split kwargs into (fields of ``parent_class``, everything else).

    "Everything else" becomes ``rot_options`` forwarded to ``cube_derotate``
    (and may hold ``algo_params``).  Reference: ``utils_param.py:133-164``.
    """
    mine, rest = {}, {}
    for key, value in initial_kwargs.items():
        (mine if hasattr(parent_class, key) or key in _ALWAYS_CLASS_KEYS else rest)[key] = value
    return mine, rest


def setup_parameters(params_obj, fkt, **add_params):
    """Keyword arguments for ``fkt`` taken from the attributes of ``params_obj``; entries of
    ``add_params`` win over attributes of the same name (``utils_param.py:60-121``)."""
    attrs = dict(vars(params_obj))
    attrs.update(add_params)
    return {name: attrs[name] for name in signature(fkt).parameters if name in attrs}
