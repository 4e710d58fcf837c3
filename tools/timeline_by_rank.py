#!/usr/bin/env python3
"""Finish time of every warp of a traced launch by residency rank of its CTA and by hardware warp slot.

    python tools/kernel_timeline.py c5r && python tools/timeline_by_rank.py gpurun_out/timeline_raw_c5r.npy

The SM's warp schedulers favour the lower hardware warp slots: with equal shares, the warps in slots 0-7 (the CTA that
became resident first) finish at 1/3 of the kernel's duration, slots 8-15 at 57 %, 16-23 at 80 %, and the last CTA runs
with 2 warps per scheduler at the end (profiles/r2_rank_finish_times.txt)."""
import numpy as np, sys
raw=np.load(sys.argv[1])
t=raw[:,:5].astype(np.float64); t0=t[:,0].min()
done=(t[:,2]-t0)*1e-3
cta=(raw[:,6]>>32).astype(int)
warpid=(raw[:,5]&0xffffffff).astype(int)
print(' total %.1f us; finish by cta//148:'%((t[:,4]-t0).max()*1e-3),[round(done[(cta//148)==r].mean(),1) for r in range(4)], ' by warpid//8:',[round(done[(warpid//8)==r].mean(),1) for r in range(4)])
# rank vs warpid consistency
print(' fraction of warps whose warpid//8 == cta//148: %.3f'%np.mean((warpid//8)==(cta//148)))
