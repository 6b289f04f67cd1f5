"""Drop-in replacements for the reference's `geometry.gshell_tets` / `geometry.hmsdf_tets_split` modules."""
from .gshell_tets import GShell_Tets  # noqa: F401
from .hmsdf_tets_split import hmSDF_Tets  # noqa: F401
