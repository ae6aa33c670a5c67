"""Literal restatement of the TDAC scoring / allocation loops (TEST INFRASTRUCTURE; scripts/calibration.py:47-90, :98-118 of the
reference, loop for loop).  Parity status: restated from the source lines -- the reference generator itself needs its samplers and a
trained diffusion model and is not run here ("parity unpinned" for this helper; the arithmetic is elementwise torch)."""
import torch
import torch.nn as nn


def scores(feature_map, dense_r=3.0):
    dense_num = torch.zeros(len(feature_map), dtype=torch.int16)
    for i in range(len(feature_map)):                                             # :49-54
        for j in range(len(feature_map)):
            if i != j:
                mse = torch.mean((feature_map[i] - feature_map[j]) ** 2)
                if mse <= dense_r:
                    dense_num[i] = dense_num[i] + 1
    cos = nn.CosineSimilarity(dim=1, eps=1e-6)
    cos_dis = torch.zeros(len(feature_map))
    for i in range(len(feature_map)):                                             # :64-68
        for j in range(len(feature_map)):
            if i != j:
                cos_dis[i] = cos_dis[i] + torch.sum(1 - cos(feature_map[i], feature_map[j]))
    return dense_num, cos_dis


def allocation(feature_map, lamda, calib_num_samples, dense_r=3.0):
    dense_num, cos_dis = scores(feature_map, dense_r)
    dense_num_normal = (dense_num - dense_num.min()) / (dense_num.max() - dense_num.min())       # :56
    cos_dis_normal = (cos_dis - cos_dis.min()) / (cos_dis.max() - cos_dis.min())                  # :70
    w = dense_num_normal + lamda * cos_dis_normal
    prob = w / torch.sum(w)
    t_num = (prob * calib_num_samples).round().to(torch.int64)                                   # :76
    t_error = calib_num_samples - torch.sum(t_num)
    _, t_num_sort = torch.sort(t_num, descending=True)
    if t_error >= 0:
        t_num[t_num_sort[:t_error]] += 1
    else:
        for i in reversed(range(len(t_num))):
            if t_error == 0:
                break
            if t_num[i] > 0:
                t_num[i] -= 1
                t_error = t_error + 1
    return t_num


def assemble(all_sample, t, num_samples):
    out = []
    for i in range(t.numel() // num_samples):                                      # :104-118
        calib_data = None
        t1 = t[i * num_samples:(i + 1) * num_samples]
        for now_rt, sample_t in enumerate(all_sample):
            if calib_data is None:
                calib_data = torch.zeros_like(sample_t)
            mask = t1 == now_rt
            if mask.any():
                calib_data += sample_t * mask.float().view(-1, *([1] * (sample_t.dim() - 1)))
        out.append(calib_data)
    return torch.cat(out)
