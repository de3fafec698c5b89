"""Multi-GPU behind the API (tsb_job_*): one process drives several contexts — here two contexts on the one GPU of the
test box (gpu_ids may repeat a device) and, when the box has more, every GPU — plus the device-side summary reduction,
the asynchronous result read-back and two contexts driven from two host threads."""
import threading

import numpy as np
import pytest
import torch

import parity_util as PU

T = PU.T
pytestmark = pytest.mark.gpu


def _single(ctx, text, n, ov, **kw):
    ckt, b, _ = PU.run_gpu(ctx, text, n, ov, out=T.OUT_STATS, **kw)
    return ckt, b


@pytest.mark.parametrize("name,n", [("rlc", 3000), ("diode2", 5001), ("transformer1", 1000)])
def test_job_over_two_contexts_equals_one_batch(ctx, name, n):
    """Contiguous shards [g*N/G, (g+1)*N/G): parameters scattered, results gathered in job order — bit-identical to one
    batch of the same instances; summary statistics reduced on each device and merged."""
    text = T.BUNDLED[name]
    ov = PU.draws(name, T.Circuit.from_netlist(text), n)
    ckt, b = _single(ctx, text, n, ov)
    card = ckt.analysis_card()
    gpus = list(range(torch.cuda.device_count())) if torch.cuda.device_count() > 1 else [0, 0, 0]
    job = T.Job(gpus, text, n)
    assert [lo for lo, _ in job.shards()] == [n * g // len(gpus) for g in range(len(gpus))] and job.shards()[-1][1] == n
    for (d, p), v in ov.items():
        job.set_param(d, p, v)
    job.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=T.OUT_STATS)
    job.sync()
    assert np.array_equal(job.status(), b.status()) and np.array_equal(job.rows(), b.rows())
    s_job, s_one = job.stats(), b.stats_all()
    assert np.array_equal(s_job, s_one, equal_nan=True)
    sm = job.summary()
    assert np.array_equal(sm["min"], np.nanmin(s_one[0], axis=1)) and np.array_equal(sm["max"], np.nanmax(s_one[1], axis=1))
    assert np.allclose(sm["sum"], s_one[2].sum(axis=1), rtol=1e-12, atol=1e-300)
    assert sm["rows"] == int(b.rows().sum()) and np.array_equal(sm["totals"], b.totals())
    # one batch's own device-side summary agrees as well
    s1 = b.summary()
    assert np.array_equal(s1["min"], sm["min"]) and np.array_equal(s1["max"], sm["max"]) and s1["rows"] == sm["rows"]


def test_job_waveform_is_routed_to_the_owning_gpu(ctx):
    text = T.BUNDLED["rc"]
    n = 77
    ov = PU.draws("rc", T.Circuit.from_netlist(text), n)
    ckt, b, _ = PU.run_gpu(ctx, text, n, ov, cap_rows=320)
    card = ckt.analysis_card()
    job = T.Job([0, 0], text, n)
    for (d, p), v in ov.items():
        job.set_param(d, p, v)
    job.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=T.OUT_WAVE, cap_rows=320)
    job.sync()
    for i in (0, 37, 38, 39, 76):
        assert np.array_equal(job.waveform(i, 320), b.waveform(i))
    with pytest.raises(T.TsbError):
        job.waveform(n, 320)


def test_async_fetch_overlaps_and_matches(ctx):
    """tsb_result_fetch_async into pinned buffers while another batch runs; a re-run of the same batch waits for the copy."""
    text = T.BUNDLED["rlc"]
    n = 4096
    ov = PU.draws("rlc", T.Circuit.from_netlist(text), n)
    ckt, b = _single(ctx, text, n, ov)
    ref = (b.stats_all(), b.rows(), b.status(), b.counters())
    card = ckt.analysis_card()
    ncol = len(ckt.columns(T.AN_TRAN))
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    st, rw, ss, cn = pin((4, ncol, n), torch.float64), pin((n,), torch.int64), pin((n,), torch.int32), pin((8, n), torch.int64)
    other = ckt.batch(n)
    for (d, p), v in ov.items():
        other.set_param(d, p, v * 1.01)
    for _ in range(3):
        b.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=T.OUT_STATS)
        b.fetch_async(st, rw, ss, cn)
        other.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=T.OUT_STATS)
        b.sync()
        assert np.array_equal(st, ref[0], equal_nan=True) and np.array_equal(rw, ref[1])
        assert np.array_equal(ss, ref[2]) and np.array_equal(cn, ref[3])
        st[:] = 0
    other.sync()


def test_two_contexts_from_two_host_threads(built):
    """What a cgo host does: one goroutine / OS thread per context, both on the same GPU, at the same time."""
    out = {}

    def work(k, name):
        c = built.Context(0)
        text = T.BUNDLED[name]
        ov = PU.draws(name, T.Circuit.from_netlist(text), 512, seed=10 + k)
        _, b, _ = PU.run_gpu(c, text, 512, ov, out=T.OUT_STATS)
        _, ores = PU.run_oracle(text, 16, {kk: v[:16] for kk, v in ov.items()}, want_wave=False, want_stats=True)
        out[k] = (b.stats_all()[:, :, :16], ores["stats"], b.counters()[:4, :16], ores["counters"][:, :4])

    th = [threading.Thread(target=work, args=(k, nm)) for k, nm in enumerate(("rc", "diode2", "rlc", "mosfet1"))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert len(out) == 4
    for k, (sg, so, cg, co) in out.items():
        so = np.transpose(so, (1, 2, 0))
        fin = np.isfinite(so)
        assert np.all(np.abs(sg[fin] - so[fin]) <= 1e-9 * np.abs(so[fin]) + 1e-12), k
        assert np.array_equal(cg.T, co), k
