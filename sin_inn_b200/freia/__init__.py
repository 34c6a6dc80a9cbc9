"""FrEIA-compatible operator API (pre-v0.2 call pattern used by the reference, archs.py:4-5,26-71)
running on the libsininn sm_100a kernels.  `framework` mirrors FrEIA.framework, `modules` mirrors
FrEIA.modules; see INTEGRATION.md for aliasing this package as `FrEIA` so the reference's
unmodified archs.py runs on these kernels."""
from . import framework, modules  # noqa: F401
