import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from fullsize import FULL, load_full
from ndpolator_b200.cndpolator import Table
path = [p for p in FULL if p.endswith('full_c3.npz')][0]
z, w, first, count = load_full(path)
T = Table(w.axes, w.grid('cuda'), len(w.basic_axes))
q = w.queries(first, count, 'cuda')
none_nan = np.unpackbits(z['none_nan'])[:count].astype(bool)
for rep in range(6):
    for m in ((0, 2, 1) if rep % 2 == 0 else (1,)):
        out = (torch.full((count, w.vdim), 12345.0, dtype=torch.float64, device='cuda'), torch.full((count, 1), 12345.0, dtype=torch.float64, device='cuda'))
        gi, gd = T.ndpolate(q, m, 0, out=out)
        gi = gi.cpu().numpy(); gd = gd.cpu().numpy()
        st = {k: T.stat(k) for k in ('offgrid', 'hard', 'ties', 'launches', 'fused')}
        if m == 1:
            bad = np.argwhere((gi[none_nan].view(np.uint64) != z['interps_m1_off'].view(np.uint64)).any(axis=1))[:, 0]
            rows = np.nonzero(none_nan)[0][bad]
            print(f'rep {rep} m {m}: {len(bad)} bad rows, sentinel rows {(gi == 12345.0).any(axis=1).sum()}, nan rows {np.isnan(gi).any(axis=1).sum()}', st)
            if len(bad):
                c, d = T.find_nearest(q, 1, 0)
                c = c.cpu().numpy()
                for r in rows[:6]:
                    print('   row', r, 'q', q[r].cpu().numpy(), 'ours', gi[r, :2], 'dist', gd[r], 'want dist', z['dists_m1'][r], 'coords ours', c[r], 'golden', z['coords_m1'][r])
        else:
            print(f'rep {rep} m {m}: sentinel rows {(gi == 12345.0).any(axis=1).sum()}', st)
