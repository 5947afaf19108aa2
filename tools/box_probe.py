"""What the GPU box offers the CPU arm and the host-buffer path: host cores and memory, whether numba imports, how fast the
staged reference VecGame (oracle/_ref, see oracle/make_ref.py) runs there, and the bare pinned-memory PCIe rates.

    python tools/box_probe.py [--out gpurun_out/box_probe_r02.json] [--max-log2 22]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--out", default="gpurun_out/box_probe_r02.json")
    p.add_argument("--max-log2", type=int, default=22)
    a = p.parse_args()
    out = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    try:
        with open("/proc/meminfo") as fh:
            out["mem_total_gb"] = int(fh.readline().split()[1]) / 1e6
        with open("/proc/cpuinfo") as fh:
            models = [ln.split(":", 1)[1].strip() for ln in fh if ln.startswith("model name")]
        out["cpu_model"] = models[0] if models else None
    except Exception as exc:  # noqa: BLE001
        out["meminfo_error"] = repr(exc)
    try:
        import numba

        out["numba"] = numba.__version__
    except Exception as exc:  # noqa: BLE001
        out["numba"] = None
        out["numba_error"] = repr(exc)

    if out["numba"]:
        import numpy as np

        from oracle import make_ref
        from oracle import oracle as orc

        out["ref_staged"] = make_ref.available()
        if out["ref_staged"]:
            gn = make_ref.import_reference()
            lib = orc.load_lib()
            numba.set_num_threads(out["affinity"])
            rows = []
            for lg in range(16, a.max_log2 + 1, 2):
                m = 1 << lg
                t0 = time.perf_counter()
                vg = gn.VecGame(m)
                vg.reset(0)
                acts = np.empty((m,), np.int64)
                row = {"games": m, "ctor_s": time.perf_counter() - t0, "prepare_s": [], "step_s": []}
                for t in range(6):
                    t0 = time.perf_counter()
                    vg.prepare()
                    t1 = time.perf_counter()
                    lib.orc_random_valid_actions(vg._data.ctypes.data, m, t, acts.ctypes.data)
                    t2 = time.perf_counter()
                    vg.step(acts)
                    t3 = time.perf_counter()
                    row["prepare_s"].append(t1 - t0)
                    row["step_s"].append(t3 - t2)
                rows.append(row)
                print(json.dumps(row), flush=True)
            out["numba_threads"] = numba.get_num_threads()
            out["numba_layer"] = numba.threading_layer()
            out["reference_vecgame"] = rows

    import torch

    if torch.cuda.is_available():
        dev = torch.device("cuda", 0)
        pcie = {}
        for name, nbytes in (("d2h_420MB", 420_000_000), ("h2d_17MB", 16_777_216), ("d2h_84MB", 84_000_000)):
            host = torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True)
            devt = torch.zeros((nbytes,), dtype=torch.uint8, device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(6):
                e0.record()
                if name.startswith("d2h"):
                    host.copy_(devt, non_blocking=True)
                else:
                    devt.copy_(host, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            pcie[name] = {"ms": best, "gbs": nbytes / best / 1e6}
        out["pcie"] = pcie
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
