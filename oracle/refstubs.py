import sys, types
def install():
    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            d = dict(d or {}); d.update(kw)
            for k, v in d.items(): self[k] = v
        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict): v = EasyDict(v)
            elif isinstance(v, (list, tuple)): v = type(v)(EasyDict(x) if isinstance(x, dict) and not isinstance(x, EasyDict) else x for x in v)
            super().__setitem__(k, v)
        __setattr__ = __setitem__
        def __getattr__(self, k):
            try: return self[k]
            except KeyError: raise AttributeError(k)
    m = types.ModuleType("easydict"); m.EasyDict = EasyDict; sys.modules["easydict"] = m
    class Dummy:
        def __init__(self,*a,**k): pass
        def __getattr__(self, k): return Dummy()
        def __call__(self,*a,**k): return Dummy()
        def __hash__(self): return id(self)
        def __iter__(self): return iter([])
    class DM(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"): raise AttributeError(k)
            return Dummy()
    for name in ["rdkit","rdkit.Chem","rdkit.RDLogger","rdkit.Chem.Draw","rdkit.Chem.rdchem","rdkit.Chem.AllChem","rdkit.Chem.Descriptors","pyemd"]:
        mm = DM(name); sys.modules[name] = mm
    sys.modules["rdkit"].__version__="0"; sys.modules["rdkit"].Chem = sys.modules["rdkit.Chem"]; sys.modules["rdkit"].RDLogger = sys.modules["rdkit.RDLogger"]
    for name in ["toponetx","toponetx.classes","toponetx.classes.combinatorial_complex"]:
        sys.modules[name] = types.ModuleType(name)
    class CombinatorialComplex: pass
    sys.modules["toponetx.classes.combinatorial_complex"].CombinatorialComplex = CombinatorialComplex
    sys.modules["toponetx.classes"].combinatorial_complex = sys.modules["toponetx.classes.combinatorial_complex"]
    sys.modules["toponetx"].classes = sys.modules["toponetx.classes"]
    sys.path.insert(0, "/root/reference")
