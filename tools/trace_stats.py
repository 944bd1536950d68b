"""Summary of a marching-kernel trace (tools/trace_probe.py)."""
import sys
import numpy as np
for f in sys.argv[1:]:
    a = np.loadtxt(f)
    t = a[:, 1:]
    print(f, "tiles", len(a), "total cycles", t.max())
    w = t[:, 1] - t[:, 0]; m = t[:, 2] - t[:, 1]; st = t[:, 3] - t[:, 2]
    print(" consumer: wait full med %.0f mean %.0f | mma med %.0f mean %.0f | store+release med %.0f mean %.0f" % (np.median(w), w.mean(), np.median(m), m.mean(), np.median(st), st.mean()))
    d = t[:, 5] - t[:, 4]; i = t[:, 6] - t[:, 5]
    print(" nbr: dep wait med %.0f mean %.0f | issue med %.0f mean %.0f" % (np.median(d), d.mean(), np.median(i), i.mean()))
    si = t[:, 8] - t[:, 7]
    print(" src: issue med %.0f mean %.0f" % (np.median(si), si.mean()))
    l = t[:, 1] - t[:, 6]; print(" nbr issued -> full med %.0f mean %.0f" % (np.median(l), l.mean()))
    l2 = t[:, 1] - t[:, 8]; print(" src issued -> full med %.0f" % np.median(l2))
    done = np.sort(t[:, 3]); print(" mean cycles between tile completions %.0f" % np.diff(done).mean())
