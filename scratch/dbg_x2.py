import sys
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np
import golden_util as gu, oracle_lib as ol, gtf_b200
fx = gu.load('barrel100_cfg1')
hb = gu.stage_batch(fx,'seed')
ob = ol.OracleBatch(hb); ob.seed(); ob.cluster(0,1.0,2.0); ob.extract(); ob.extrapolate_stage(2.0)
b = gtf_b200.EventBatch(ob.hb)
lab_g = b.CCA(); lab_o = ob.cca().copy()
d = np.nonzero(lab_g!=lab_o)[0]
print('label diff nodes', d, 'gpu', lab_g[d], 'oracle', lab_o[d], 'alive', ob.hb['alive'][d], 'sub', ob.hb['sub'][d], 'sub_state', ob.hb['sub_state'][ob.hb['sub'][d]])
print('S', ob.S, 'sub_state', ob.hb['sub_state'], 'sub_off', ob.hb['sub_off'])
# chain on GPU
ob2 = ol.OracleBatch(hb); ob2.seed()
b2 = gtf_b200.EventBatch(hb); b2.seed()
ALL = ("alive","active","merged","tse","uts","degree","edge_w")
def cmp(tag):
    print(tag, gu.compare_states(b2.download(), ob2.hb, ALL, rtol=1e-7))
cmp('seed')
ob2.cluster(0,1.0,2.0); b2.cluster(0,1.0,2.0); cmp('c1')
r_o = ob2.extract(); r_g = b2.extract(); print('x1 acc eq', np.array_equal(r_o[1], r_g[1]), r_o[0], r_g[0]); cmp('x1')
ob2.extrapolate_stage(2.0); b2.extrapolate_stage(2.0); cmp('e2')
lo = ob2.cca().copy(); lg = b2.CCA(); dd = np.nonzero(lo!=lg)[0]; print('e2 label diffs', dd, lo[dd], lg[dd])
r_o = ob2.extract(); r_g = b2.extract(); print('x2 acc eq', np.array_equal(r_o[1], r_g[1]), r_o[0], r_g[0]); cmp('x2')
dd = np.nonzero(r_o[1]!=r_g[1])[0]
print('acc diff nodes', dd, 'labels o', lo[dd], 'g', lg[dd])
for r in np.unique(np.concatenate([lo[dd], lg[dd]])):
    mo = np.nonzero(lo==r)[0]; mg = np.nonzero(lg==r)[0]
    print(' root', r, 'oracle members', mo, 'gpu members', mg, 'pv o', r_o[2][r], r_o[3][r], 'pv g', r_g[2][r], r_g[3][r])
