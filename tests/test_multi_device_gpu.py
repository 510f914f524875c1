"""GPU: boundary hardening (round-1 verdict weak #9, ADVICE functional.py:17).
The reference's demos place models with device_map="auto" (pipeline/demos/demo_models.py:37) and its serving worker
calls generate() from a worker thread (pipeline/serve/model_worker.py:246): ops must follow their tensors' device
(per-device function attributes / SM count / stream) and the library must be safe to call from several threads."""
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(dev, seed):
    """Module + inputs; built on the caller's thread (torch's global RNG is not per-thread)."""
    from otter_b200.modeling_otter import OtterGatedCrossAttentionBlock
    torch.manual_seed(seed)
    gb = OtterGatedCrossAttentionBlock(dim=256, dim_visual=128).to(dev)
    with torch.no_grad():
        gb.attn_gate.fill_(0.5), gb.ff_gate.fill_(0.5)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(2, 40, 256, generator=g).to(dev)
    media = torch.randn(2, 1, 64, 128, generator=g).to(dev)
    loc = torch.zeros(2, 40, dtype=torch.bool, device=dev)
    loc[:, 0] = True
    return gb, x, media, loc


def _run(gb, x, media, loc):
    gb.zero_grad()
    x = x.detach().clone().requires_grad_(True)
    y = gb(x, media, media_locations=loc)
    y.float().pow(2).mean().backward()
    return y.detach().float().cpu(), x.grad.float().cpu(), gb.feed_forward[1].weight.grad.float().cpu()


def _block_step(dev, seed):
    return _run(*_make(dev, seed))[:2]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """Tensors on cuda:1 while cuda:0 is current: same results as on cuda:0, nothing launched on the wrong device."""
    y0, g0 = _block_step("cuda:0", 3)
    assert torch.cuda.current_device() == 0
    y1, g1 = _block_step("cuda:1", 3)                       # current device stays 0: the ops must switch by themselves
    assert torch.cuda.current_device() == 0
    assert torch.equal(y0, y1) and torch.equal(g0, g1)
    from otter_b200 import functional as F
    from otter_b200._lib import OtbError
    a = torch.randn(8, 64, device="cuda:0").to(torch.bfloat16)
    w = torch.randn(16, 64, device="cuda:1").to(torch.bfloat16)
    with pytest.raises(OtbError, match="different devices"):
        F.linear_fwd(a, w)


def test_concurrent_threads_on_their_own_streams():
    """Two Python threads, each on its own CUDA stream, running forward+backward concurrently (ctypes releases the GIL):
    results equal the single-threaded run bit for bit; the descriptor cache sees hits."""
    from otter_b200 import _lib
    jobs = [_make("cuda:0", s) for s in (5, 6)]
    want = [_run(*j) for j in jobs]
    torch.cuda.synchronize()
    got, errs = [None, None], []

    def work(i):
        try:
            st = torch.cuda.Stream(device="cuda:0")
            st.wait_stream(torch.cuda.default_stream())
            with torch.cuda.stream(st):
                for _ in range(4):
                    got[i] = _run(*jobs[i])
                st.synchronize()
        except Exception as e:                              # surfaced below
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for a, b in zip(got, want):
        assert all(torch.equal(u, v) for u, v in zip(a, b))
    lib = _lib.load()
    assert lib.otb_tmap_cache_stat(0) > 0 and lib.otb_tmap_cache_stat(1) > 0
