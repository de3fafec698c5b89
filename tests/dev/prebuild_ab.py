"""Development aid: pre-build (nvcc, here) the kernels tests/gpu_ab.py will ask for — one deck, several $TSB_EXTRA_DEFINES
variants, every launch-bounds candidate — so that no GPU-box time goes into NVRTC.
Usage: python tests/dev/prebuild_ab.py <deck> "<opts k=v,..>|<defines>" ..."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G  # noqa: E402
import parity_util as PU  # noqa: E402

T = PU.T
deck = sys.argv[1]
jobs = []
for spec in sys.argv[2:] or [""]:
    o, _, rest = spec.partition("|")
    defs, _, _ = rest.partition("|")
    kw = {}
    for item in filter(None, o.split(",")):
        k, v = item.split("=")
        kw[k] = int(v)
    if defs:
        os.environ["TSB_EXTRA_DEFINES"] = defs
    else:
        os.environ.pop("TSB_EXTRA_DEFINES", None)
    ckt = T.Circuit.from_netlist(T.BUNDLED[deck])
    b = ckt.batch(2)
    for (d, p), v in PU.draws(deck, ckt, 2).items():
        b.set_param(d, p, v)
    strict = kw.get("strict_fp", 0)
    for mb in range(1, 7):
        opts = T.default_opts(**{**kw, "strict_fp": strict, "min_blocks": mb})
        jobs.append((b.kernel_source(opts), b.kernel_key(opts), "false" if strict else "true", f"{deck} [{spec}] mb{mb}"))
os.environ.pop("TSB_EXTRA_DEFINES", None)
with ThreadPoolExecutor(max_workers=8) as ex:
    for label, r in zip((j[3] for j in jobs), ex.map(lambda j: G._nvcc_cubin(j[0], j[1], j[2]), jobs)):
        print(label, r, flush=True)
