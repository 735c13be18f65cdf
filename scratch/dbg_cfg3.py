import sys
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np
import golden_util as gu, oracle_lib as ol, gtf_b200
from gtf_b200 import synth
hb = synth.event_to_host(synth.barrel_event(10000, seed=3300), 0); hb.pop('truth'); hb.pop('orig_id')
ob = ol.OracleBatch(hb); ob.seed()
b = gtf_b200.EventBatch(hb); b.seed()
ALL = ("alive","active","merged","tse","uts","degree","edge_w")
print('seed', gu.compare_states(b.download(), ob.hb, ALL))
ob.cluster(0,1.0,2.0); b.cluster(0,1.0,2.0)
g = b.download()
print('c1', gu.compare_states(g, ob.hb, ALL))
d = np.nonzero(g['active']!=ob.hb['active'])[0]
print('active diffs', len(d), d[:10])
hm = np.nonzero(g['has_merged']!=ob.hb['has_merged'])[0]; print('has_merged diffs', len(hm), hm[:10])
for i in hm[:5]:
    s0,s1 = hb['in_off'][i], hb['in_off'][i+1]
    print(' node', i, 'deg', s1-s0, 'gpu', g['has_merged'][i], 'oracle', ob.hb['has_merged'][i])
m = (g['has_merged']>0)&(ob.hb['has_merged']>0)
for f in ('m_a','m_b','m_c','m_p00','m_p01','m_p11','m_p22','m_prior'):
    a,bb = g[f][m], ob.hb[f][m]; rel = np.abs(a-bb)/np.maximum(np.abs(a),np.abs(bb)); rel[np.isnan(rel)]=0
    k = np.argmax(rel); print(f, rel[k], a[k], bb[k], (rel>1e-9).sum())
