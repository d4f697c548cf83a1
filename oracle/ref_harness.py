"""Import the UNMODIFIED reference from /root/reference on CPU torch (build container only).

TEST INFRASTRUCTURE.  Applies the monkey-patches listed in SURVEY.md section 8c without
touching the reference sources:
  (1) ``.cuda()`` is hard-coded all over the reference (model/transfer.py:317,325,
      368-385,466-468,703-705) -> identity when no GPU is present;
  (2) ``torch.load(args.pre_model)`` unpickles a whole module (model/transfer.py:323)
      -> ``weights_only=False``;
  (3) ``evaluation2.test_model`` returns a numpy scalar on modern numpy and the caller
      does ``ndcg.cpu()`` (model/transfer.py:813) -> wrap to return a tensor;
  (4) ``--numworkers 0``.
  (5) on CUDA: TF32 off for cuDNN / matmul so that conv1 / conv2 are fp32 like the CPU path, and
      ``MFbasemode.test`` hands its per-batch NDCG back on the host (``np.array([cuda tensors])``
      in evalution/evaluation2.py:24-25 refuses device tensors).
Search order for the tree: $SML_REFERENCE, /root/reference (build container), baseline/_ref (the
staged copy that travels to the GPU box, oracle/stage_reference.py).  Only bench.py's reference /
cpu_baseline legs (through oracle/ref_arm.py, in a subprocess) and oracle/gen_golden.py import this
file; nothing under sml_b200/ does.
"""
from __future__ import annotations

import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find():
    for c in (os.environ.get("SML_REFERENCE"), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")):
        if c and os.path.isdir(os.path.join(c, "model")):
            return c
    return os.environ.get("SML_REFERENCE", "/root/reference")


REF = _find()


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "model"))


def load(device=None):
    """Returns a namespace with the reference modules (MF, conv_transfer, transfer,
    evaluation2, evaluation, evalution_function, dataset, dataset2, main_yelp).
    ``device``: "cpu" forces the CPU path even when a GPU is present (shim 1), "cuda" runs the
    reference's own torch-CUDA path (shim 5); default: cpu unless CUDA is available."""
    import types
    import torch
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    if device == "cpu":
        torch.Tensor.cuda = lambda self, *a, **k: self           # shim (1)
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.manual_seed = lambda *a, **k: None
    else:
        torch.backends.cudnn.allow_tf32 = False                  # shim (5)
        torch.backends.cuda.matmul.allow_tf32 = False
    _orig_load = torch.load

    def _load(*a, **k):                                          # shim (2)
        k.setdefault("weights_only", False)
        return _orig_load(*a, **k)
    torch.load = _load

    import importlib
    ns = types.SimpleNamespace()
    ns.MF = importlib.import_module("model.MF")
    ns.conv_transfer = importlib.import_module("model.conv_transfer")
    ns.dataset = importlib.import_module("data.dataset")
    ns.dataset2 = importlib.import_module("data.dataset2")
    ns.evaluation2 = importlib.import_module("evalution.evaluation2")
    ns.evaluation = importlib.import_module("evalution.evaluation")
    ns.evalution_function = importlib.import_module("evalution.evalution_function")
    ns.transfer = importlib.import_module("model.transfer")
    ns.main_yelp = importlib.import_module("main_yelp")
    ns.main_news = importlib.import_module("main_news")

    if device != "cpu":                                          # shim (5): per-batch NDCG back on the host
        _test = ns.MF.MFbasemode.test

        def _test_host(self, *a, **k):
            h, n, idx = _test(self, *a, **k)
            return h, (n.cpu() if torch.is_tensor(n) else n), idx
        ns.MF.MFbasemode.test = _test_host

    _tm = ns.evaluation2.test_model

    def _test_model(*a, **k):                                    # shim (3)
        r, n = _tm(*a, **k)
        return r, torch.as_tensor(float(n))
    ns.transfer.test_model = _test_model
    ns.raw_test_model = _tm
    return ns


def theta_to_numpy(net):
    """one_transfer module -> dict of numpy arrays keyed like its state_dict."""
    return {k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}
