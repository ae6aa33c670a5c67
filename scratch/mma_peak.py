import ctypes, os, torch
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmmapeak.so"))
lib.run_peak.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
dev = torch.device("cuda:0")
cyc = torch.zeros(148, dtype=torch.int64, device=dev)
for (n, ce, rd) in [(192, 0, 0), (192, 1, 0), (192, 2, 0), (192, 4, 0), (192, 0, 4), (192, 1, 4), (256, 1, 0), (256, 1, 4), (64, 1, 0)]:
    iters, grid, f16 = 4096, 148, 0
    st = torch.cuda.current_stream().cuda_stream
    lib.run_peak(n, iters, f16, grid, cyc.data_ptr(), st, ce, rd); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = lib.run_peak(n, iters, f16, grid, cyc.data_ptr(), st, ce, rd); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ops = 2.0 * 128 * n * 32 * 4 * iters * grid
    print(f"i8 N={n:3d} commit every {ce} k-steps (4 MMAs each), tmem readers {rd}: rc={rc} {cyc[0].item() / (4 * iters):7.1f} cycles/MMA  {ops / (ms * 1e-3) / 1e12:8.1f} TOP/s", flush=True)
