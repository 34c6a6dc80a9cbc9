"""ORACLE ONLY. Sequential-chain subset of FrEIA.framework (pre-v0.2).

Call sites restated: /root/reference/archs.py:26 (InputNode(c,h,w,name=)),
:28-31,35-38,61-68 (Node(prev, module_type, module_args, name=)), :70
(OutputNode(prev, name=)), :71 (ReversibleGraphNet(nodes, verbose=False)).
The graph the reference builds is a pure chain, so only chains are supported.
"""
import torch.nn as nn


class Node:
    def __init__(self, inputs, module_type, module_args, conditions=None, name=None):
        if isinstance(inputs, (list, tuple)):
            assert len(inputs) == 1, "oracle shim supports chains only"
            inputs = inputs[0]
        if isinstance(inputs, tuple):
            inputs = inputs[0]
        self.input = inputs
        self.module_type = module_type
        self.module_args = module_args
        self.name = name
        self.module = None
        self.output_dims = None

    def build(self):
        dims_in = [self.input.output_dims[0]]
        self.module = self.module_type(dims_in, **self.module_args)
        self.output_dims = self.module.output_dims(dims_in)


class InputNode(Node):
    def __init__(self, *dims, name="node"):
        self.input = None
        self.name = name
        self.module = None            # pre-v0.2: input/output nodes own no module,
        self.output_dims = [tuple(dims)]  # but still occupy a ModuleList slot

    def build(self):
        pass


class OutputNode(Node):
    def __init__(self, inputs, name="node"):
        if isinstance(inputs, (list, tuple)):
            inputs = inputs[0]
        self.input = inputs
        self.name = name
        self.module = None
        self.output_dims = None

    def build(self):
        self.output_dims = self.input.output_dims


class ReversibleGraphNet(nn.Module):
    """Runs node modules in list order (reversed with rev=True), each as
    ``module([x], rev=rev)[0]``; returns the bare tensor (pre-v0.2: no log-det
    in the return value -- the reference slices the result directly,
    lit_wrapper.py:45-46)."""

    def __init__(self, node_list, ind_in=None, ind_out=None, verbose=True):
        super().__init__()
        self.node_list = node_list
        for n in node_list:
            n.build()
        # ModuleList over *all* nodes so that key index == node index
        # (state_dict keys like ``module_list.3.s1.0.weight``).
        self.module_list = nn.ModuleList([n.module for n in node_list])

    def forward(self, x, c=None, rev=False):
        mods = [m for m in self.module_list if m is not None]
        if rev:
            mods = mods[::-1]
        for m in mods:
            x = m([x], rev=rev)[0]
        return x
