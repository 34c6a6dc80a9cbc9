"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.

CPU restatement of the slice of FrEIA (github.com/VLL-HD/FrEIA, pre-v0.2 API,
version un-pinned by the reference: it is absent from requirements.txt:1-8 and
not vendored) that /root/reference/archs.py:4-5,26-71 uses.  It exists so the
reference's own ``archs.py`` can be imported *unmodified* to pin the oracle, and
so the CPU baseline can run.  PARITY AT THE FrEIA BOUNDARY IS UNPINNED: FrEIA is
not installable offline and the reference has no tests;
the semantics here follow the published pre-v0.2 source and are self-checked
(exact inverse, autograd consistency, known channel orders).
"""
from . import framework, modules  # noqa: F401
