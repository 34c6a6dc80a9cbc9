"""FrEIA.framework drop-in (chain graphs): InputNode / Node / OutputNode / ReversibleGraphNet.

Replaces the graph container the reference instantiates at archs.py:26-71.  The node list is compiled
once into an engine plan; ``net(x)`` / ``net(x, rev=True)`` then run fused on the libsininn kernels and
return ONE NCHW-contiguous fp32 tensor (pre-v0.2 behaviour the callers rely on, lit_wrapper.py:45-46)."""
import torch.nn as nn

from .. import engine as E
from . import modules as Fm


class Node:
    def __init__(self, inputs, module_type, module_args, conditions=None, name=None):
        if conditions:
            raise E.SininnError("conditional nodes are not supported")
        while isinstance(inputs, (list, tuple)):
            if len(inputs) != 1 and not (len(inputs) == 2 and isinstance(inputs[1], int)):
                raise E.SininnError("only single-input chain graphs are supported")
            inputs = inputs[0]
        self.input, self.module_type, self.module_args, self.name = inputs, module_type, module_args, name
        self.module, self.output_dims = None, None

    def build(self):
        dims_in = [self.input.output_dims[0]]
        self.module = self.module_type(dims_in, **self.module_args)
        self.output_dims = self.module.output_dims(dims_in)


class InputNode(Node):
    def __init__(self, *dims, name="node"):
        self.input, self.name, self.module = None, name, None
        self.output_dims = [tuple(dims)]

    def build(self):
        pass


class OutputNode(Node):
    def __init__(self, inputs, name="node"):
        while isinstance(inputs, (list, tuple)):
            inputs = inputs[0]
        self.input, self.name, self.module, self.output_dims = inputs, name, None, None

    def build(self):
        self.output_dims = self.input.output_dims


def op_from_module(m):
    """Plan op for a FrEIA-protocol module or an IRN block."""
    if hasattr(m, "_op"):
        return m._op()
    raise E.SininnError(f"no sm_100a kernel path for module type {type(m).__name__}")


class ReversibleGraphNet(nn.Module):
    def __init__(self, node_list, ind_in=None, ind_out=None, verbose=True):
        super().__init__()
        self.node_list = node_list
        for i, n in enumerate(node_list):
            if i > 0 and n.input is not node_list[i - 1]:
                raise E.SininnError("only chain graphs are supported (each node must consume the previous one)")
            n.build()
        # every node owns a slot so that key index == node index ("module_list.3.s1.0.weight")
        self.module_list = nn.ModuleList([n.module for n in node_list])
        self.in_dims = node_list[0].output_dims[0]
        self.engine_config = None
        self._plan = None

    def plan(self):
        if self._plan is None:
            ops = [op_from_module(m) for m in self.module_list if m is not None]
            self._plan = E.Plan(ops, self.in_dims)
        return self._plan

    def forward(self, x, c=None, rev=False):
        if c is not None:
            raise E.SininnError("conditional inputs are not supported")
        cfg = self.engine_config or E.default_config()
        return E.run_network(self.plan(), x, rev, cfg)
