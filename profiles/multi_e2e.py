"""Product-level multi-GPU entry point (intfft_multi_exec_host): ONE pinned host batch, sharded inside the library over
the visible devices, against (a) one device alone and (b) the bare concurrent H2D + D2H copies of the same bytes on all
devices at once (the ceiling this box allows).  One JSON line per device count."""
import sys, os, json, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import intfftk_b200 as ib

ndev = torch.cuda.device_count()
g = ib.Generics(NFFT=12, DATA_WIDTH=16, FORMAT=0)
per_dev = 65536                                           # c2 batch per device
cudart = torch.cuda.cudart()
for nd in [n for n in (1, 2, 4, 8) if n <= ndev]:
    batch = per_dev * nd
    hin = ib.HostBuffer((batch, 4096, 2), np.int16)
    hout = ib.HostBuffer((batch, 4096, 2), np.int16)
    hin.array[...] = 1
    hin.array[:64] = np.arange(64 * 8192, dtype=np.int16).reshape(64, 4096, 2)
    m = ib.Multi(g, batch, 0, list(range(nd)))
    m.exec_host_ptr(hin.ptr, hout.ptr)                    # warm-up: staging rings, streams
    steps = 3
    t0 = time.perf_counter()
    for _ in range(steps):
        m.exec_host_ptr(hin.ptr, hout.ptr)
    ms = (time.perf_counter() - t0) / steps * 1e3
    m.close()
    # bare copies: per device one H2D and one D2H stream moving that device's shard, all devices at once
    bufs, streams = [], []
    shard_bytes = per_dev * 4096 * 2 * 2
    for d in range(nd):
        with torch.cuda.device(d):
            bufs.append((torch.empty(shard_bytes, dtype=torch.uint8, device=f"cuda:{d}"), torch.empty(shard_bytes, dtype=torch.uint8, device=f"cuda:{d}")))
            streams.append((torch.cuda.Stream(d), torch.cuda.Stream(d)))
    h_in_t = torch.from_numpy(hin.array.view(np.uint8).reshape(-1))
    h_out_t = torch.from_numpy(hout.array.view(np.uint8).reshape(-1))
    def copies():
        for d in range(nd):
            with torch.cuda.device(d):
                with torch.cuda.stream(streams[d][0]):
                    bufs[d][0].copy_(h_in_t[d * shard_bytes:(d + 1) * shard_bytes], non_blocking=True)
                with torch.cuda.stream(streams[d][1]):
                    h_out_t[d * shard_bytes:(d + 1) * shard_bytes].copy_(bufs[d][1], non_blocking=True)
        for d in range(nd):
            torch.cuda.synchronize(d)
    copies()
    t0 = time.perf_counter()
    for _ in range(steps):
        copies()
    cms = (time.perf_counter() - t0) / steps * 1e3
    samples = batch * 4096
    print(json.dumps({"devices": nd, "frames": batch, "e2e_ms": round(ms, 2), "e2e_gsamples_s": round(samples / ms / 1e6, 2),
                      "bare_copy_ms": round(cms, 2), "copy_ceiling_gsamples_s": round(samples / cms / 1e6, 2),
                      "frac_of_ceiling": round(cms / ms, 3),
                      "pinned": "intfft_host_alloc (cudaHostAllocPortable)"}), flush=True)
    del bufs, streams
    hin.close(); hout.close()
