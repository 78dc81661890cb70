"""Analyse a k_linsolve_tb job-time dump (EQ_LSX_JOBTIMES=file): u64[G*NBP][4] = start, first chunk ready, end, smid.
usage: jobtimes_tb.py file NBP"""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4).astype(np.int64)
NBP = int(sys.argv[2]); G = a.shape[0] // NBP
a = a.reshape(G, NBP, 4)
t0 = a[:, :, 0][a[:, :, 0] > 0].min()
st, rd, en, sm = (a[:, :, 0] - t0) / 1e3, (a[:, :, 1] - t0) / 1e3, (a[:, :, 2] - t0) / 1e3, a[:, :, 3]
dur = en - st
print(f"G={G} NBP={NBP} total {en.max():.0f} us; job duration us: mean {dur.mean():.0f} min {dur.min():.0f} p50 {np.median(dur):.0f} p90 {np.percentile(dur,90):.0f} max {dur.max():.0f}")
print(f"wait for first chunk us: mean {(rd-st).mean():.0f} p50 {np.median(rd-st):.0f} p90 {np.percentile(rd-st,90):.0f} max {(rd-st).max():.0f}")
print("sum of job durations / (CTAs*total):", dur.sum() / (740 * en.max()))
for g in (0, 1, G // 2, G - 1):
    print(f"group {g}: start of band 0 {st[g,0]:.0f}, band NBP/2 {st[g,NBP//2]:.0f}, last {st[g,-1]:.0f}; end last {en[g,-1]:.0f}; "
          f"mean start spacing between bands {np.diff(st[g]).mean():.1f} us; mean dur {dur[g].mean():.0f}")
lag = st[:, 1:] - st[:, :-1]
print(f"start lag to band above us: p10 {np.percentile(lag,10):.1f} p50 {np.median(lag):.1f} p90 {np.percentile(lag,90):.1f}")
endlag = en[:, 1:] - en[:, :-1]
print(f"end lag to band above us: p10 {np.percentile(endlag,10):.1f} p50 {np.median(endlag):.1f} p90 {np.percentile(endlag,90):.1f}")
glag = en[1:, :-1] - en[:-1, 1:]
print(f"end(b,g+1) - end(b+1,g) us: p10 {np.percentile(glag,10):.1f} p50 {np.median(glag):.1f} p90 {np.percentile(glag,90):.1f}")
# per-band slope: how much later does band b end than band b-1, averaged over bands 50..NBP-50 of group 0 and last
for g in (0, G - 1):
    print(f"group {g}: end-time slope {(en[g,-50]-en[g,50])/(NBP-100):.2f} us/band")
