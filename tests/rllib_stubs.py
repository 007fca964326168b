"""Stand-ins for the third-party imports of the reference's RLlib adapter (/root/reference/src/gcm/ray_gcm.py:1-18):
`gym`, `ray.rllib.*` and `torch_geometric` are not installable here.  Only what the adapter touches at
import / construction / forward time is provided; torch_geometric resolves to this repository's gcm.nn layers, so the
adapter's `from gcm.gcm import DenseGCM, ...` and its DenseGCM calls run on the product package unchanged."""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ("/root/reference/src/gcm/ray_gcm.py", os.path.join(ROOT, "baseline", "_ref", "gcm", "ray_gcm.py"))


def reference_adapter_path():
    for p in CANDIDATES:
        if os.path.exists(p):
            return p
    return None


class Space:
    def __init__(self, dim):
        self.dim = dim


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install():
    """Put the stand-in modules into sys.modules; returns the names added (for removal)."""
    import gcm.nn as gnn

    class TorchModelV2:
        def __init__(self, obs_space, action_space, num_outputs, model_config, name):
            self.obs_space, self.action_space = obs_space, action_space
            self.view_requirements = {}

    class SlimFC(torch.nn.Module):
        def __init__(self, in_size, out_size, activation_fn=None, initializer=None):
            super().__init__()
            self._model = torch.nn.Linear(in_size, out_size)
            if initializer is not None:
                initializer(self._model.weight)

        def forward(self, x):
            return self._model(x)

    def normc_initializer(std=1.0):
        def init(t):
            with torch.no_grad():
                t.normal_(0, 1)
                t *= std / torch.sqrt(t.pow(2).sum(1, keepdim=True))
        return init

    mods = {
        "gym": _mod("gym", spaces=_mod("gym.spaces", Space=Space, utils=_mod("gym.spaces.utils", flatdim=lambda s: s.dim))),
        "ray": _mod("ray"),
        "ray.rllib": _mod("ray.rllib"),
        "ray.rllib.models": _mod("ray.rllib.models"),
        "ray.rllib.models.torch": _mod("ray.rllib.models.torch"),
        "ray.rllib.models.torch.torch_modelv2": _mod("ray.rllib.models.torch.torch_modelv2", TorchModelV2=TorchModelV2),
        "ray.rllib.models.torch.fcnet": _mod("ray.rllib.models.torch.fcnet", FullyConnectedNetwork=object),
        "ray.rllib.models.torch.misc": _mod("ray.rllib.models.torch.misc", SlimFC=SlimFC, normc_initializer=normc_initializer),
        "ray.rllib.models.torch.recurrent_net": _mod("ray.rllib.models.torch.recurrent_net", RecurrentNetwork=object),
        "ray.rllib.utils": _mod("ray.rllib.utils"),
        "ray.rllib.utils.typing": _mod("ray.rllib.utils.typing", ModelConfigDict=dict, TensorType=torch.Tensor),
        "ray.rllib.utils.torch_utils": _mod("ray.rllib.utils.torch_utils", one_hot=None),
        "ray.rllib.policy": _mod("ray.rllib.policy"),
        "ray.rllib.policy.sample_batch": _mod("ray.rllib.policy.sample_batch", SampleBatch=dict),
        "ray.rllib.policy.view_requirement": _mod("ray.rllib.policy.view_requirement", ViewRequirement=object),
        "torch_geometric": _mod("torch_geometric", nn=gnn),
        "torch_geometric.nn": gnn,
        "torch_geometric.data": _mod("torch_geometric.data", Data=object, Batch=object),
    }
    mods["gym.spaces"] = mods["gym"].spaces
    mods["gym.spaces.utils"] = mods["gym"].spaces.utils
    added = [k for k in mods if k not in sys.modules]
    for k in added:
        sys.modules[k] = mods[k]
    return added


def load_reference_adapter():
    """The reference's ray_gcm.py, executed from where it lies, on top of the stand-ins and of THIS package's gcm."""
    path = reference_adapter_path()
    added = install()
    try:
        spec = importlib.util.spec_from_file_location("reference_ray_gcm", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k in added:
            sys.modules.pop(k, None)
    return mod
