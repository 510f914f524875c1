import os, time, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
a = torch.randn(2048, 2048); b = torch.randn(2048, 2048)
for t in (8, 16, 32, 64, 128):
    torch.set_num_threads(t); a @ b
    t0 = time.perf_counter()
    for _ in range(3): a @ b
    dt = (time.perf_counter() - t0) / 3
    print(f"threads {t:4d}: {2 * 2048 ** 3 / dt / 1e9:8.1f} GFLOP/s")
