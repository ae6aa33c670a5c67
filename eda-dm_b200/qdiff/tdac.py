"""TDAC -- timestep-aligned calibration-set allocation (the step in front of the hot path).

The reference's generators (scripts/calibration.py: TDAC_cifar_calib_data_generator :12-154 and its LDM twins) collect the input of
one mid-block module at every sampling step, then score the T steps by
  * density  (:47-56): how many other steps' features lie within mean-squared distance `dense_r` (3.0),
  * variety  (:58-68): the summed (1 - cosine similarity over channels) to all other steps,
and hand out `calib_num_samples` samples over the steps in proportion to density + lamda * variety (:70-90), finally assembling
the calibration tensor from the stored trajectory (:104-118).  The two scores are O(T^2) loops of elementwise passes over
feature tensors of several MB in the reference.  Here both are Gram matrices: one [T, D] x [D, T] product each (fp64 on the
device, D = samples * channels * pixels), ||a - b||^2 = ||a||^2 + ||b||^2 - 2 a.b and sum cos = <a / |a|, b / |b|>, so T^2 tensor
passes become two matmuls; the allocation arithmetic is kept operation for operation.  The sampler loops that produce the
trajectories stay with the caller (out of scope).
"""
import torch


def tdac_scores(feature_map, dense_r: float = 3.0, eps: float = 1e-6):
    """feature_map: sequence of T tensors [N, C, ...] (or one [T, N, C, ...] tensor).  Returns (dense_num int16 [T], cos_dis fp32 [T])
    as scripts/calibration.py:47-68 computes them."""
    f = torch.stack(list(feature_map)) if not torch.is_tensor(feature_map) else feature_map
    T = f.shape[0]
    f64 = f.reshape(T, -1).double()
    D = f64.shape[1]
    gram = f64 @ f64.t()
    sq = torch.diagonal(gram)
    mse = (sq[:, None] + sq[None, :] - 2.0 * gram) / D
    off = ~torch.eye(T, dtype=torch.bool, device=f.device)
    dense_num = ((mse <= dense_r) & off).sum(1).to(torch.int16)
    # cosine over dim 1 of [N, C, ...]: normalise every (sample, position) channel vector, then the sum of cosines is a dot product
    g = f.reshape(T, f.shape[1], f.shape[2], -1).double()                       # [T, N, C, P]
    norm = g.norm(dim=2, keepdim=True).clamp_min(eps)
    gn = (g / norm).reshape(T, -1)
    cos_sum = gn @ gn.t()
    positions = g.shape[1] * g.shape[3]
    cos_dis = ((positions - cos_sum) * off).sum(1).float()
    return dense_num.cpu(), cos_dis.cpu()


def tdac_allocation(feature_map, lamda: float, calib_num_samples: int, dense_r: float = 3.0):
    """samples per timestep, int64 [T] summing to calib_num_samples (scripts/calibration.py:55-90)"""
    dense_num, cos_dis = tdac_scores(feature_map, dense_r)
    dense_norm = (dense_num - dense_num.min()) / (dense_num.max() - dense_num.min())
    cos_norm = (cos_dis - cos_dis.min()) / (cos_dis.max() - cos_dis.min())
    w = dense_norm + lamda * cos_norm
    prob = w / torch.sum(w)
    t_num = (prob * calib_num_samples).round().to(torch.int64)
    t_error = int(calib_num_samples - torch.sum(t_num))
    _, order = torch.sort(t_num, descending=True)
    if t_error >= 0:
        t_num[order[:t_error]] += 1
    else:
        for i in reversed(range(len(t_num))):
            if t_error == 0:
                break
            if t_num[i] > 0:
                t_num[i] -= 1
                t_error += 1
    assert int(torch.sum(t_num)) == calib_num_samples
    return t_num


def tdac_assemble(all_sample, t_num, num_samples: int, seq=None, generator=None):
    """Calibration tensor and timesteps from a stored trajectory (scripts/calibration.py:98-150): `all_sample[k]` is the batch
    [num_samples, ...] at sampling step k; sample i of every chunk is taken from the step its (shuffled) allocation names.
    Returns (calib_data [calib_num_samples, ...], step index per sample, timestep per sample when `seq` is given)."""
    device = all_sample[0].device
    t = torch.cat([torch.full((int(n),), k, dtype=torch.long) for k, n in enumerate(t_num)])
    t = t[torch.randperm(t.numel(), generator=generator)].to(device)
    traj = torch.stack(list(all_sample))                                       # [T, num_samples, ...]
    chunks = []
    for i in range(t.numel() // num_samples):
        t1 = t[i * num_samples:(i + 1) * num_samples]
        chunks.append(traj[t1, torch.arange(num_samples, device=device)])       # one gather instead of T masked accumulations
    calib = torch.cat(chunks)
    if seq is None:
        return calib, t, None
    seq_t = torch.as_tensor(list(seq), device=device)
    return calib, t, seq_t[(len(seq) - 1) - t]
