"""Device-side timeline of one SMC step inside the device-resident loop (block-0 %globaltimer stamps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modppl_b200 as m
from bench import observations

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
scheme = m.SYSTEMATIC_NESTED
ys = observations(64)
ps = m.ParticleSystem(m.lgssm4(), 1 << log2n, seed=1, dtype="f32")
ps.upload_observations(ys)
ps.run(0, 20, scheme)
ms = ps.run(20, 30, scheme)
tr = ps.device_trace()
t0 = tr[13]
names = {13: "extend block 0 past its wait", 14: "extend last block done", 9: "section pass start", 5: "section pass: last block in", 7: "section pass: top level done",
         10: "level-1 pass start", 11: "expansion block 0 past its wait"}
print(f"ms/step {ms / 30:.4f}")
for k, v in sorted(names.items(), key=lambda kv: tr[kv[0]]):
    print(f"  {v:40s} {(tr[k] - t0) / 1e3:8.1f} us")
