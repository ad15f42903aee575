"""Reading a saved episode (the reference's `.arl` files: pickle.dump of a list of Environment.save_state() copies,
main.py:136-147, environment.py:36-40) WITHOUT the classes that wrote it: every class of the `environment` /
`generator` packages is replaced by a stub that keeps the pickled attribute dict.  The same reader therefore opens
files written by the reference and by the drop-in layer, and the two can be compared attribute by attribute."""
import io
import pickle

import numpy as np


class Stub:
    """Stand-in for a pickled class instance: `cls` is 'module.ClassName', `state` the pickled attribute dict."""
    cls = "?"

    def __setstate__(self, state):
        self.__dict__.update(state)


class NeutralUnpickler(pickle.Unpickler):
    PACKAGES = ("environment", "generator", "utils", "agents")

    def find_class(self, module, name):
        if module.split(".")[0] in self.PACKAGES:
            return type(name, (Stub,), {"cls": module + "." + name})
        return super().find_class(module, name)


def load_neutral(blob_or_path):
    if isinstance(blob_or_path, (bytes, bytearray)):
        return NeutralUnpickler(io.BytesIO(blob_or_path)).load()
    with open(blob_or_path, "rb") as f:
        return NeutralUnpickler(f).load()


def describe(obj):
    """{attribute: (kind, dtype, shape)} of a stub, for structure comparisons."""
    out = {}
    for k, v in vars(obj).items():
        if isinstance(v, np.ndarray):
            out[k] = ("array", str(v.dtype), tuple(v.shape))
        elif isinstance(v, Stub):
            out[k] = ("object", v.cls, ())
        else:
            out[k] = (type(v).__name__, "", ())
    return out


# the attributes gui/visualize.py reads from each object of a state (visualize.py:77-95, 189-248)
VIEWER_READS = {
    "Walls": ("map",),
    "AntsVisualization": ("ants", "holding", "reward_state", "mandibles"),
    "CircleObstaclesVisualization": ("centers", "radiuses"),
    "PheromoneVisualization": ("phero", "max_val", "color"),
    "FoodVisualization": ("qte",),
    "RLVisualization": ("heatmap",),
    "AnthillVisualization": ("x", "y", "radius", "food"),
}


def compare_states(got, ref, rtol=1e-5, atol=1e-7):
    """A state (stub Environment) written by the code under test against the reference's: same object classes in the
    same order, same attribute names / dtypes / shapes, integer arrays identical, float arrays within the parity bar."""
    assert got.cls == ref.cls == "environment.environment.Environment"
    assert (got.w, got.h, got.max_time, got.timestep) == (ref.w, ref.h, ref.max_time, ref.timestep)
    assert [o.cls for o in got.objects] == [o.cls for o in ref.objects]
    for a, b in zip(got.objects, ref.objects):
        name = a.cls.rsplit(".", 1)[1]
        if name == "Walls":                 # the live object (walls.py:16-17); its environment is compared separately
            assert np.array_equal(a.map, b.map) and a.map.dtype == b.map.dtype
            continue
        da, db = describe(a), describe(b)
        assert da == db, (name, da, db)
        for k in VIEWER_READS[name]:
            assert k in da, (name, k)
        for k, v in vars(b).items():
            w = getattr(a, k)
            if isinstance(v, np.ndarray):
                if v.dtype.kind == "f":
                    np.testing.assert_allclose(w, v, rtol=rtol, atol=atol, err_msg="%s.%s" % (name, k))
                else:
                    assert np.array_equal(w, v), "%s.%s" % (name, k)
            elif not isinstance(v, Stub):
                assert w == v, (name, k, w, v)
