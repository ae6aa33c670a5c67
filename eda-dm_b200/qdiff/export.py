"""Packed W4A8 checkpoint export / import (the step behind the hot path).

The reference publishes W4A8 weights (README.md:66-73, model sizes in assets/compression.png) but ships no serializer: nothing in
scripts/ or qdiff/ calls torch.save on a quantized model.  `export_packed` writes what a deployment needs and nothing else:
per QuantModule the integer weight codes exactly as the integer path uses them (hard AdaRound rounding after reconstruction,
adaptive_rounding.py:50-58; nearest rounding before it, quant_layer.py:267-268), nibble-packed two per byte for <= 4 bit, with the
per-channel (delta, zero_point), the activation quantizers' (delta, zero_point, bits), the attention quantizers, and the
remaining FP tensors (norms, biases, embeddings) in fp32.  `load_packed` restores a freshly built QuantModel from it: weights
come back as (code - zp) * delta -- which the plain nearest-rounding quantizer maps to the same codes again, so no AdaRound state
is needed -- and every forward is bit-identical to the exporting model's.
"""
import io

import torch

from edadm import ops
from .quant_layer import QuantModule, UniformAffineQuantizer

FORMAT = "edadm-packed-v1"


def _pack_nibbles(codes: torch.Tensor) -> torch.Tensor:
    flat = codes.reshape(codes.shape[0], -1)
    if flat.shape[1] % 2:
        flat = torch.cat([flat, torch.zeros_like(flat[:, :1])], 1)
    return (flat[:, 0::2] | (flat[:, 1::2] << 4)).contiguous()


def _unpack_nibbles(packed: torch.Tensor, k: int) -> torch.Tensor:
    out = torch.stack([packed & 0xF, packed >> 4], dim=2).reshape(packed.shape[0], -1)
    return out[:, :k]


def _q_entry(q):
    return None if q is None or q.delta is None else {
        "delta": q.delta.detach().float().cpu().clone(), "zero_point": torch.as_tensor(q.zero_point).detach().float().cpu().clone(),
        "n_bits": int(q.n_bits), "inited": bool(q.inited)}


def export_packed(qnn, path=None):
    """-> dict (also written with torch.save when `path` is given).  Needs a CUDA model: codes come from edadm_pack_weight."""
    blob = {"format": FORMAT, "layers": {}, "quantizers": {}, "fp": {}}
    owned = set()
    for name, m in qnn.named_modules():
        if not isinstance(m, QuantModule):
            continue
        wqs, aqs = m._quantizers()
        w = m.weight.detach()
        w4 = w.reshape(w.shape[0], w.shape[1], *([1] * (4 - w.dim()))) if w.dim() < 4 else w
        if w.dim() == 3:
            w4 = w.unsqueeze(2)
        bounds = [(0, w.shape[1])] if m.split == 0 else [(0, m.split), (m.split, w.shape[1])]
        parts = []
        for wq, (c0, c1) in zip(wqs, bounds):
            pw = ops.pack_weight(w4, wq.delta, wq.zero_point, wq.n_levels, alpha=getattr(wq, "alpha", None), c_begin=c0, c_end=c1,
                                 want_codes=True, w4=False)
            codes = pw.codes.cpu()
            parts.append({"codes": _pack_nibbles(codes) if wq.n_bits <= 4 else codes.reshape(codes.shape[0], -1).contiguous(),
                          "k": int(codes[0].numel()), "shape": tuple(codes.shape), "n_bits": int(wq.n_bits),
                          "delta": wq.delta.detach().float().reshape(-1).cpu().clone(),
                          "zero_point": torch.as_tensor(wq.zero_point).detach().float().reshape(-1).cpu().clone()})
        blob["layers"][name] = {"weights": parts, "split": int(m.split), "weight_shape": tuple(w.shape),
                                "bias": None if m.bias is None else m.bias.detach().float().cpu().clone(),
                                "act": [_q_entry(a) for a in aqs], "disable_act_quant": bool(m.disable_act_quant)}
        owned.update(id(p) for p in m.parameters(recurse=True))
    for name, q in qnn.named_modules():
        if isinstance(q, UniformAffineQuantizer) and ".weight_quantizer" not in name and not name.endswith("act_quantizer") \
                and not name.endswith("act_quantizer_0") and q.delta is not None:
            blob["quantizers"][name] = _q_entry(q)            # attention q / k / v / softmax quantizers
    layer_prefixes = tuple(n + "." for n in blob["layers"])
    for name, p in list(qnn.named_parameters()) + list(qnn.named_buffers()):
        if id(p) in owned or name.startswith(layer_prefixes) or "quantizer" in name:
            continue
        blob["fp"][name] = p.detach().float().cpu().clone()
    if path is not None:
        torch.save(blob, path)
    return blob


def packed_nbytes(blob) -> int:
    n = 0
    stack = [blob]
    while stack:
        o = stack.pop()
        if torch.is_tensor(o):
            n += o.numel() * o.element_size()
        elif isinstance(o, dict):
            stack.extend(o.values())
        elif isinstance(o, (list, tuple)):
            stack.extend(o)
    return n


def load_packed(qnn, blob_or_path):
    """Restore a freshly built QuantModel (same architecture, `set_first_last_layer_to_8bit` / split flags as at export) in place."""
    blob = torch.load(blob_or_path, map_location="cpu") if not isinstance(blob_or_path, dict) else blob_or_path
    if blob.get("format") != FORMAT:
        raise ValueError("not an edadm packed checkpoint")
    dev = next(qnn.parameters()).device
    named = dict(qnn.named_modules())

    def restore_q(q, e):
        if e is None:
            return
        q.bitwidth_refactor(e["n_bits"])
        q.zero_point = e["zero_point"].to(dev)
        d = e["delta"].to(dev)
        q.delta = torch.nn.Parameter(d) if q.leaf_param else d
        q.inited = e["inited"]

    for name, e in blob["layers"].items():
        m = named[name]
        if e["split"] and m.split == 0:
            m.split = e["split"]
            m.set_split()
        deq, wqs = [], []
        for part in e["weights"]:
            codes = (_unpack_nibbles(part["codes"], part["k"]) if part["n_bits"] <= 4 else part["codes"]).reshape(part["shape"]).float()
            shape = [-1] + [1] * (len(part["shape"]) - 1)
            deq.append((codes - part["zero_point"].reshape(shape)) * part["delta"].reshape(shape))
            wq = UniformAffineQuantizer(**m.weight_quant_params)
            wq.bitwidth_refactor(part["n_bits"])
            wshape = [-1] + [1] * (len(e["weight_shape"]) - 1)
            wq.delta, wq.zero_point, wq.inited = part["delta"].reshape(wshape).to(dev), part["zero_point"].reshape(wshape).to(dev), True
            wqs.append(wq)
        w = torch.cat(deq, 1).reshape(e["weight_shape"]).to(dev)
        with torch.no_grad():
            m.weight.copy_(w)
            m.org_weight = w.clone()
            if e["bias"] is not None:
                m.bias.copy_(e["bias"].to(dev))
                m.org_bias = e["bias"].to(dev).clone()
        m.weight_quantizer = wqs[0]
        if len(wqs) > 1:
            m.weight_quantizer_0 = wqs[1]
        _, aqs = m._quantizers()
        for a, ea in zip(aqs, e["act"]):
            restore_q(a, ea)
        m.disable_act_quant = e["disable_act_quant"]
        m._packed = None
    for name, e in blob["quantizers"].items():
        restore_q(named[name], e)
    params = dict(qnn.named_parameters())
    params.update(dict(qnn.named_buffers()))
    with torch.no_grad():
        for name, t in blob["fp"].items():
            params[name].copy_(t.to(dev))
    return qnn
