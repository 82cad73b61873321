"""In-situ timeline of ONE decode step (CUDA-graph replay, PDL) from a -DTGIS_STEP_TIMELINE build:
    python scripts/build_variant.py stl -DTGIS_STEP_TIMELINE
    TGIS_ENGINE_LIB=$PWD/vllm_tgis_adapter_b200/lib/libtgis_engine_stl.so python scripts/step_timeline.py [layers] [batch] [prompt]
CTA 0 / thread 0 of every launch stamps %globaltimer at entry, after griddepcontrol.wait and at exit.  Printed per kernel
of the last decode step: start (us from the step's first entry), time waiting for its dependency, body time, and the gap
between the previous kernel's exit and this kernel's dependency-wait return (= exposed boundary latency)."""
import ctypes as C
import dataclasses
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vllm_tgis_adapter_b200.engine import _lib  # noqa: E402
from vllm_tgis_adapter_b200.engine.core import PRESETS, NativeEngine, make_sampling_params  # noqa: E402
from vllm_tgis_adapter_b200.engine.loader import load_synthetic_weights, rope_cos_sin  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
P = int(sys.argv[3]) if len(sys.argv) > 3 else 512
lib = _lib.load_library()
rc = lib.tgis_k_step_timeline_enable()
assert rc == 0, f"not a -DTGIS_STEP_TIMELINE build (rc={rc})"
mc = dataclasses.replace(PRESETS["llama3-8b"], n_layers=L, max_model_len=1024)
# one prefill step for all prompts: every sequence then decodes in lock step and the LAST steps of the run (the ones the
# 4096-record ring still holds) are full-batch steps, not the tail of a staggered job
eng = NativeEngine(mc, max_num_seqs=B, max_batched_tokens=B * P, kv_cache_bytes=(B * (P + 96) * 2 * L * 8 * 128 * 2 * 5) // 4)
load_synthetic_weights(eng, mc, 0, 0)
eng.load_weight("tgis.rope_cos_sin", rope_cos_sin(mc))
rs = np.random.RandomState(0)
prompts = [rs.randint(1000, mc.vocab - 1000, size=P).tolist() for _ in range(B)]
G = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sp = make_sampling_params(greedy=True, max_tokens=G, min_tokens=G)
for i, p in enumerate(prompts):
    eng.add_request(f"r{i}", p, sp)
eng.run_until_idle()
st = eng.status()
buf = np.zeros(8 + 4096 * 8, dtype=np.uint64)
assert lib.tgis_k_step_timeline_read(buf.ctypes.data_as(C.POINTER(C.c_uint64))) == 0
n = int(buf[0])
recs = buf[8:8 + 4096 * 4].reshape(4096, 4)
extra = buf[8 + 4096 * 4:].reshape(4096, 4)   # GEMM records: first MMA, last TMA issued, last accumulator ready, rstd ready
idx = [(i % 4096) for i in range(max(0, n - 4096), n)]
# word 0 = kernel id | N << 8 in the low 32 bits; bits 32..51 = ns between the last exit stamp of ANY warp of the earlier
# kernels and this kernel's dependency wait returning (0 when the kernel has no wait mark)
rows = [(int(recs[i][0]) & 0xffffffff, int(recs[i][1]), int(recs[i][2]), int(recs[i][3]), (int(recs[i][0]) >> 32) & 0xfffff,
         tuple(int(x) for x in extra[i]))
        for i in idx if recs[i][3] != 0]
rows.sort(key=lambda r: r[1])
NAMES = {1: "norm", 2: "gemm", 3: "attn_decode", 4: "attn_merge", 5: "attn_prefill", 6: "sampler", 8: "gather", 9: "ar_norm",
         20: "STEP_BEGIN", 21: "META_LANDED", 22: "STEP_END"}
# last decode step = everything after the second-to-last sampler record
samp = [i for i, r in enumerate(rows) if (r[0] & 0xff) == 6]
lo = samp[-2] + 1 if len(samp) >= 2 else 0
step = [r for r in rows[lo:samp[-1] + 1] if (r[0] & 0xff) not in (20, 21, 22)]
t0 = step[0][1]
# step-to-step decomposition from the marker kernels (20 = before the metadata copy, 21 = after it, 22 = after the result
# copy) of the last decode steps
begins = [r for r in rows if (r[0] & 0xff) == 20]
landed = [r for r in rows if (r[0] & 0xff) == 21]
ends = [r for r in rows if (r[0] & 0xff) == 22]
sampx = [rows[i] for i in samp]
nm = min(len(begins), len(landed), len(ends))
begins, landed, ends = begins[len(begins) - nm:], landed[len(landed) - nm:], ends[len(ends) - nm:]
if nm > 4:
    K = min(20, nm - 1)
    dec = []
    for j in range(len(begins) - K, len(begins)):
        b, l, e, pe = begins[j], landed[j], ends[j], ends[j - 1]
        sx = [x for x in sampx if b[1] < x[1] < e[1]]
        first = next((r for r in rows if r[1] > l[3] and (r[0] & 0xff) not in (20, 21, 22)), None)
        dec.append(((b[1] - pe[3]) / 1e3, (l[3] - b[1]) / 1e3, ((first[2] or first[1]) - l[3]) / 1e3 if first else 0.0,
                    (e[3] - sx[-1][3]) / 1e3 if sx else 0.0, (e[3] - pe[3]) / 1e3))
    d = np.array(dec)
    print(f"last {K} steps (us): prev STEP_END -> STEP_BEGIN (host turnaround + launch latency) mean {d[:, 0].mean():.1f} "
          f"min {d[:, 0].min():.1f} max {d[:, 0].max():.1f} | metadata copy {d[:, 1].mean():.1f} | -> first kernel running "
          f"{d[:, 2].mean():.1f} | sampler exit -> STEP_END (result copy) {d[:, 3].mean():.1f} | period {d[:, 4].mean():.1f}")
    print("  per-step turnaround:", " ".join(f"{x:.0f}" for x in d[:, 0]))
    print("  per-step BEGIN->END :", " ".join(f"{x - y:.0f}" for x, y in zip(d[:, 4], d[:, 0])))
print(f"decode ms/step (engine events) = {st.gpu_decode_ms / max(st.decode_steps, 1):.4f}; graph launches {st.graph_launches}; "
      f"{len(step)} kernel launches in the last step, span {(step[-1][3] - t0) / 1e3:.1f} us")
print(f"{'kernel':>14} {'N':>7} {'start':>8} {'wait':>7} {'body':>7} {'gap_prev_exit->waited':>22}  = prev tail + dep latency")
prev_exit = None
agg = {}
gph = {}
out_rows = []
for kid, te, tw, tx, lat, ex in step:
    name = NAMES.get(kid & 0xff, str(kid & 0xff))
    Nn = kid >> 8
    tw_ = tw if tw else te
    gap = (tw_ - prev_exit) / 1e3 if prev_exit else 0.0
    key = f"{name}:{Nn}" if Nn else name
    a = agg.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0])
    if name == "gemm" and ex[0] and ex[2]:   # phases of the instrumented CTA relative to its dependency wait returning
        g_ = gph.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0])
        g_[0] += 1
        g_[1] += (ex[0] - tw_) / 1e3          # -> first MMA issued
        g_[2] += (ex[1] - tw_) / 1e3          # -> last TMA issued
        g_[3] += (ex[2] - tw_) / 1e3          # -> last accumulator ready
        g_[4] += ((ex[3] - tw_) / 1e3) if ex[3] else 0.0   # -> rstd ready (fused-norm consumers)
    a[0] += 1
    a[1] += (tw_ - te) / 1e3
    a[2] += (tx - tw_) / 1e3
    a[3] += gap
    a[4] += lat / 1e3
    out_rows.append({"kernel": key, "start_us": (te - t0) / 1e3, "wait_us": (tw_ - te) / 1e3, "body_us": (tx - tw_) / 1e3,
                     "gap_us": gap, "dep_latency_us": lat / 1e3})
    prev_exit = tx
for r in out_rows[: 2 + 8 * 2]:   # the first two layers in detail
    print(f"{r['kernel']:>14} {'':>7} {r['start_us']:8.1f} {r['wait_us']:7.1f} {r['body_us']:7.1f} {r['gap_us']:22.1f}"
          f"  = {r['gap_us'] - r['dep_latency_us']:5.1f} + {r['dep_latency_us']:4.1f}")
print("---- per kernel type over the whole step: launches, mean wait / body / gap (us) [gap = tail of the PREVIOUS kernel "
      "(its last warp exit - its CTA 0 exit) + dependency latency], sum of body+gap (us)")
tot = 0.0
for k, (c, w, b_, g, dl) in sorted(agg.items(), key=lambda kv: -(kv[1][2] + kv[1][3])):
    print(f"{k:>18} x{c:3d}  wait {w / c:6.2f}  body {b_ / c:6.2f}  gap {g / c:6.2f} [= {(g - dl) / c:5.2f} + {dl / c:4.2f}]"
          f"   sum {b_ + g:8.1f}")
    tot += b_ + g
print(f"sum of body+gap = {tot:.1f} us")
print("---- GEMM CTA 0, mean us after its dependency wait returned: first MMA issued / last TMA issued / last accumulator ready "
      "/ exit / (rstd ready)")
for k, (c, a1, a2, a3, a4) in sorted(gph.items()):
    print(f"{k:>18}  first_mma {a1 / c:6.2f}  last_tma {a2 / c:6.2f}  acc_ready {a3 / c:6.2f}  exit {agg[k][2] / agg[k][0]:6.2f}"
          f"  rstd {a4 / c:5.2f}")
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/step_timeline.json").write_text(json.dumps({"layers": L, "batch": B, "prompt": P, "rows": out_rows, "agg": agg}))
eng.close()
