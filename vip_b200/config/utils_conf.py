"""Input validation helper (``vip_hci/config/utils_conf.py:309-378``)."""
import numpy as np


def check_array(input_array, dim, msg=None):
    """Raise ``TypeError`` unless ``input_array`` is an ndarray whose ndim matches ``dim``
    (an int 1..4 or one of the tuples (1,2), (2,3), (3,4), (2,3,4)); 1-d also accepts lists/tuples."""
    if not isinstance(input_array, (list, tuple, np.ndarray)):
        raise TypeError("`input_array` must be a list, tuple of numpy ndarray")
    name = "Input array" if msg is None else "`" + msg + "`"
    valid = "`dim` must be: 1, 2, 3, 4, (1,2), (2,3), (3,4) or (2,3,4)"
    if isinstance(dim, int):
        if not 1 <= dim <= 4:
            raise ValueError(valid)
        dims, label = (dim,), str(dim)
    elif isinstance(dim, tuple):
        if dim not in ((1, 2), (2, 3), (3, 4), (2, 3, 4)):
            raise ValueError(valid)
        dims = dim
        label = ", ".join(str(d) for d in dim[:-1]) + " or " + str(dim[-1])
    else:
        raise ValueError(valid)
    if dim == 1 and isinstance(input_array, (list, tuple)):
        input_array = np.array(input_array)
    if not isinstance(input_array, np.ndarray) or input_array.ndim not in dims:
        kind = "list, tuple or a 1" if dim == 1 else label
        raise TypeError(name + " must be a " + kind + "d numpy ndarray")


sep = "―" * 80          # section separator of the reference's console output (``config/utils_conf.py``)
