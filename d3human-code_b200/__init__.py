"""d3human-code_b200: B200-native G-Shell / mSDF marching-tetrahedra extraction (the one hot path of D3-Human).

Drop-in classes (same constructor, call signature and return tuple as the reference):
    geometry.gshell_tets.GShell_Tets        <- reference geometry/gshell_tets.py:89,253
    geometry.hmsdf_tets_split.hmSDF_Tets    <- reference geometry/hmsdf_tets_split.py:89,254
"""
__version__ = "0.1.0"
