import os, sys, json
import numpy as np
sys.path.insert(0, '/root/repo')
from oracle import ref_loader as rl, oracle_np as onp
rl.load_full()
cmp = sys.modules["gcMapExplorer.lib.ccmap"]
OUT='/root/repo/tests/golden'
def case(name, A, level, method, bno=None):
    c = cmp.CCMAP(); c.gen_from_matrix(A, 1000, xlabel='chrS', ylabel='chrS')
    if bno is not None: c.bNoData = bno
    o = cmp.downSampleCCMap(c, level=level, method=method)
    o.make_readable()
    M = np.array(o.matrix)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), input=A, method='DS', kwargs=json.dumps(dict(level=level, method=method)),
                        matrix=M, bNoData=(np.asarray(o.bNoData, dtype=bool) if o.bNoData is not None else np.zeros(0, dtype=bool)),
                        bNoData_in=(bno if bno is not None else np.zeros(0, dtype=bool)),
                        minvalue=np.float64(o.minvalue), maxvalue=np.float64(o.maxvalue), binsize=np.int64(o.binsize))
    ref = onp.downsample_2d(A, level, method)
    print(name, A.shape, '->', M.shape, 'oracle bit-exact:', np.array_equal(ref, M), 'bNoData ok:', (bno is None) or np.array_equal(onp.downsample_nodata(bno, M.shape[0], level), o.bNoData))
A97 = onp.synth_map(97, seed=97)
vc = onp.normalize_vc(A97)["matrix"]          # non-integer values
A131 = onp.synth_map(131, seed=131, scale=30.0, alpha=1.3, empty_frac=0.03)
b131 = ~onp.get_nonzeros_index(A131)
for lvl in (2, 3, 4):                          # 96 % 2,3,4 == 0 -> no padding (fp32 arithmetic)
    for m in ('sum', 'mean', 'max'):
        case(f'ds_s97_l{lvl}_{m}', A97, lvl, m)
case('ds_s97vc_l4_mean', vc, 4, 'mean'); case('ds_s97vc_l12_sum', vc, 12, 'sum'); case('ds_s97vc_l16_mean', vc, 16, 'mean')
for lvl in (2, 3, 7, 10):                      # 130 % 3,7 != 0 -> padded (fp64 arithmetic); 130 % 2,10 == 0
    for m in ('sum', 'mean', 'max'):
        case(f'ds_sp131_l{lvl}_{m}', A131, lvl, m, b131)
case('ds_s97vc_l5_sum', vc, 5, 'sum'); case('ds_s97vc_l5_mean', vc, 5, 'mean'); case('ds_s97vc_l9_max', vc, 9, 'max')
