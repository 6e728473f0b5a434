"""Populations pickled by the reference (`got.savePopulation`, gori_tools.py:252-257) -> the B200
façade.

A reference pickle names its classes as `geneticalgorithm.GeneticalGorithm`, `geneticalgorithm.
Organism`, `mappy.Mappy` and `pathmaker.PathMaker`; unpickling it normally needs the reference's
modules (and their IPython / matplotlib imports) on the path.  `load_population` resolves those four
names to attribute bags instead, so a legacy file opens anywhere; `load` then hands the bag to
`GeneticalGorithm.from_population`, which packs the tables from the pickled `Mappy` / `PathMaker`
state (one copy - the reference pickles a deep copy per parent), re-evaluates the
parents on the GPU and continues from the stored generation number."""
import gzip
import io
import pickle

class _Bag(object):
    """State of a pickled reference object (its instance __dict__), no behaviour."""
    _legacy_class = None


class GeneticalGorithm(_Bag):
    _legacy_class = "geneticalgorithm.GeneticalGorithm"


class Organism(_Bag):
    _legacy_class = "geneticalgorithm.Organism"


class Mappy(_Bag):
    _legacy_class = "mappy.Mappy"


class PathMaker(_Bag):
    _legacy_class = "pathmaker.PathMaker"


_LEGACY = {("geneticalgorithm", "GeneticalGorithm"): GeneticalGorithm,
           ("geneticalgorithm", "Organism"): Organism,
           ("mappy", "Mappy"): Mappy, ("pathmaker", "PathMaker"): PathMaker}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _LEGACY:
            return _LEGACY[(module, name)]
        return super().find_class(module, name)


def load_population(path_or_file):
    """Unpickle a reference population without the reference's modules.  Returns an object with
    the reference's attributes (`_gen_parent`, `_gen_parent_fit`, `_gen_num`, `_G_sz`, ...); each
    parent has `_dna`, `_obj_val`, `_mappy`, `_pather`, ...  `.gz` files are opened transparently."""
    if hasattr(path_or_file, "read"):
        return _Unpickler(path_or_file).load()
    with open(path_or_file, "rb") as f:
        head = f.read(2)
    opener = gzip.open if head == b"\x1f\x8b" else open
    with opener(path_or_file, "rb") as f:
        return _Unpickler(io.BytesIO(f.read())).load()


def load(path_or_file, device=0, seed=None, gen_params=None):
    """Legacy pickle -> genetic_coverage_planning_b200.geneticalgorithm.GeneticalGorithm."""
    from .geneticalgorithm import GeneticalGorithm
    old = load_population(path_or_file)
    return GeneticalGorithm.from_population(old, gen_params=gen_params, device=device, seed=seed)
