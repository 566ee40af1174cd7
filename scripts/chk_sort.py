import sys, numpy as np
sys.path.insert(0, ".")
from sparkl_b200 import abi, scenes
from sparkl_b200.pipeline import context_from_scene
sc = scenes.sand3(48, 24, 24)
ctx = context_from_scene(sc)
p1 = abi.SolverParameters(stop_after_one_substep=1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 125):
    ctx.step(p1, sc.gravity)
ctx.frame_begin(p1)
ctx.stage_sort()
hdr = ctx.grid_headers(1)
ids = ctx.read(abi.ATTR_SORTED_IDS)
pos = ctx.read(abi.ATTR_POSITION)["point"]
h = np.float32(sc.cell_width)
cell = np.round(pos / h).astype(np.int64) - 2
cib = (cell & 3)
cidx = cib[:, 0] + 4 * cib[:, 1] + 16 * cib[:, 2]
nb = []; nsorted = 0; tot = 0
for hh in hdr:
    f, n = int(hh["first_particle"]), int(hh["num_particles"])
    if n == 0: continue
    c = cidx[ids[f:f + n]]
    tot += 1
    nsorted += int((np.diff(c) >= 0).all())
    for k in range(0, n, 256):
        nb.append((len(np.unique(c[k:k + 256])), min(256, n - k)))
nb = np.array(nb)
print("blocks with particles", tot, "cell-sorted", nsorted, "chunks", len(nb), "mean distinct cells/chunk", nb[:, 0].mean(), "mean particles/chunk", nb[:, 1].mean())
print("particles per cell histogram:", np.bincount(np.bincount((cell[:,0]+100) + 1000*(cell[:,1]+100) + 1000000*(cell[:,2]+100)))[:20])
