"""small end-to-end run (every entry point once; usable under a memory checker): seed, cluster, three committed iterations (with wide nodes), components,
extraction, per-stage call after iterations, download."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import gtf_b200
from gtf_b200 import synth
hbs = [synth.event_to_host(synth.barrel_event(120, seed=900 + i, eta_max=1.0 if i else 0.5, target_degree=6.0 + 6 * i), i) for i in range(3)]
hb = synth.concat_host_batches(hbs)
hb.pop("truth"); hb.pop("orig_id")
b = gtf_b200.EventBatch(hb)
b.seed(); b.cluster("track_state_estimates", 1.0, 2.0)
st = b.iterate(max_iter=3, stop_when_converged=False, record_chi2=True)
b.iterate_dry()
lab = b.CCA()
b.extract()
b.iterate(max_iter=1, stop_when_converged=False)
b.remove_state_metadata()
b.iterate(max_iter=2)
out = b.download()
print("ok", st[-1], int((out["active"] == 1).sum()), len(np.unique(lab)))
