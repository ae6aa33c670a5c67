"""Import shims for running the UNMODIFIED reference (``/root/reference``) on CPU.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` in the build container to
generate the committed fixtures under ``tests/golden/``.  Nothing on the product path, in
``bench.py`` or in the ``-m gpu`` tests imports this file: ``/root/reference`` does not exist on
the GPU box.

What is shimmed:
  * ``matplotlib``  -- imported but never used by qdiff/{block,layer,attn_layer}_recon.py:11
  * ``omegaconf.listconfig.ListConfig`` -- only touched in UNetModel.__init__ when context_dim
    is given (ldm/modules/diffusionmodules/openaimodel.py:508-512)
  * hard-coded ``'cuda'`` device strings (qdiff/block_recon.py:128, set_quantize_params.py:35)
    -- ``Tensor.cuda`` / ``Tensor.to('cuda')`` are mapped to no-ops on a CPU-only host.
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def install(reference_root: str = REFERENCE_ROOT):
    import torch

    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt

    # qdiff_control/coco_prompt.py imports dataset tooling that is irrelevant to the hot path
    for name in ("pycocotools", "pycocotools.coco", "skimage", "skimage.io"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod = types.ModuleType(name)
                if name == "pycocotools.coco":
                    mod.COCO = object
                sys.modules[name] = mod
    try:
        import omegaconf  # noqa: F401
    except Exception:
        oc = types.ModuleType("omegaconf")
        lc = types.ModuleType("omegaconf.listconfig")

        class ListConfig(list):
            pass

        lc.ListConfig = ListConfig
        oc.listconfig = lc
        sys.modules["omegaconf"] = oc
        sys.modules["omegaconf.listconfig"] = lc

    if not torch.cuda.is_available() and not getattr(torch.Tensor, "_edadm_cpu_shim", False):
        _orig_to = torch.Tensor.to

        def _to(self, *args, **kwargs):
            args = list(args)
            for i, a in enumerate(args):
                if isinstance(a, str) and a.startswith("cuda"):
                    args[i] = "cpu"
                elif isinstance(a, torch.device) and a.type == "cuda":
                    args[i] = torch.device("cpu")
            dev = kwargs.get("device")
            if isinstance(dev, str) and dev.startswith("cuda"):
                kwargs["device"] = "cpu"
            return _orig_to(self, *args, **kwargs)

        torch.Tensor.to = _to
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.empty_cache = lambda: None
        torch.Tensor._edadm_cpu_shim = True
