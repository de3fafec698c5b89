"""Development aid: pre-build (nvcc, here) the cooperative-mapping kernels gpu_ladder_perf.py will ask for, TSB_COOP_MIN_BLOCKS
variants included, so that no GPU-box time goes into NVRTC.  Usage: python tests/dev/prebuild_coop_variants.py 12,16,24 2,3,4"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G  # noqa: E402
import parity_util as PU  # noqa: E402
import random_decks as _RD
rc_ladder = getattr(_RD, os.environ.get("LADDER_KIND", "rc_ladder"))       # e.g. LADDER_KIND=diode_rc_ladder  # noqa: E402

T = PU.T
secs = [int(x) for x in sys.argv[1].split(",")]
mbs = sys.argv[2].split(",") if len(sys.argv) > 2 else [""]
jobs = []
for s in secs:
    ckt = T.Circuit.from_netlist(rc_ladder(s))
    for parts in (0, 2, 4, 8):
        if parts and ckt.coop_info(parts) is None:
            continue
        for mb in (mbs if parts else [""]):
            if mb:
                os.environ["TSB_EXTRA_DEFINES"] = f"TSB_COOP_MIN_BLOCKS={mb}"
            else:
                os.environ.pop("TSB_EXTRA_DEFINES", None)
            b = ckt.batch(2)
            for (d, p), v in PU.draws("ladder", ckt, 2, seed=5).items():
                b.set_param(d, p, v)
            if parts:
                o = T.default_opts(strict_fp=0, min_blocks=2, coop_parts=parts)
                jobs.append((b.kernel_source(o), b.kernel_key(o), f"ladder{s}:coop{parts}:mb{mb}"))
            else:
                for m in range(1, 7):       # every launch-bounds candidate the autotuner may ask for
                    o = T.default_opts(strict_fp=0, min_blocks=m)
                    jobs.append((b.kernel_source(o), b.kernel_key(o), f"ladder{s}:thread:mb{m}"))
os.environ.pop("TSB_EXTRA_DEFINES", None)


def one(j):
    r = G._nvcc_cubin(j[0], j[1], "true")
    return j[2], r


with ThreadPoolExecutor(max_workers=8) as ex:
    for label, r in ex.map(one, jobs):
        print(label, r, flush=True)
