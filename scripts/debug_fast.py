import sys; sys.path.insert(0, '.')
import numpy as np
from ov2slam_b200 import api, synth
from oracle import image_ref as R
ctx = api.Context(0)
w, h = 640, 480
for cs in (50, 16):
    im = synth.make_frame(30, w, h)
    pyr = api.Pyramid(ctx, 1, w, h, 0); pyr.build(im[None])
    fe = api.FeatureExtractor(ctx, nfast_th=10)
    cells = fe.debug_fast_cells(pyr, 0, cs, 10)
    nw = w // cs
    bad = 0
    for ci, got in enumerate(cells):
        r, c = ci // nw, ci % nw
        x, y = c * cs, r * cs
        if not (x + cs < w - 1 and y + cs < h - 1):
            assert got == []; continue
        ref = [k for k in R.fast_detect_ref(im[y:y+cs, x:x+cs], 10) if k[0] % 4 in (2, 3)]
        if got != ref:
            bad += 1
            if bad <= 4: print('cell', ci, x, y, 'got', got[:8], 'ref', ref[:8], len(got), len(ref))
    print('cs', cs, 'bad cells', bad, 'of', len(cells))
    pyr.close()
