"""Development tool: statistics of a per-tile timeline written by the -DPBTE_TRACE build (tools/trace_probe.py).
usage: timeline_stats.py <timeline file> [ring slots = 8] [loader warps per role = 2]
Columns of the file: tile, then clock stamps (cycles from the first stamp) 0 consumer starts waiting, 1 slot full, 2 MMAs done,
3 tile published, 4 neighbour loader has the free slot, 5 prerequisites stored, 6 neighbour copies issued, 7 source loader has the free
slot, 8 source copies issued, 9 descriptor read, 10 tile stored."""
import sys

import numpy as np

a = np.loadtxt(sys.argv[1])
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
t = a[:, 1:]
n = len(a)
c0, c1, c2, c3, c4, c5, c6, c7, c8, c9, c10 = [t[:, k] for k in range(11)]
print(f"tiles {n}, cycles of the slot {c3.max():.0f} = {c3.max() / n:.0f} per tile")
print("consumer, cycles per tile: wait for a full slot %.0f | descriptor %.0f | matrix blocks %.0f | wall terms, store, publish %.0f | sum %.0f"
      % ((c1 - c0).mean(), (c9 - c1).mean(), (c2 - c9).mean(), (c3 - c2).mean(), (c3 - c0).mean()))
i = np.arange(R, n)
free = c2[i - R]  # the consumer of the slot's previous tile is through its matrix blocks (it hands the slot back right after the wall terms)
print("chain from 'slot free' to 'slot full': %.0f = neighbour loader takes the slot %.0f + prerequisites %.0f + copies issued %.0f + landed %.0f"
      % ((c1[i] - free).mean(), (c4[i] - free).mean(), (c5 - c4)[i].mean(), (c6 - c5)[i].mean(), (c1 - c6)[i].mean()))
print("  the neighbour copies are issued after the source copies in %.0f %% of the tiles" % (100 * (c6 > c8).mean()))
prev = c6[i - K]
print("  the neighbour loader is still busy with its previous tile when the slot frees in %.0f %% of the tiles" % (100 * (prev > free).mean()))
for w in range(K):
    idx = np.arange(w, n, K)
    print("  neighbour warp %d: %.0f cycles between its tiles, %.0f of them from slot taken to copies issued; source warp %d: %.0f, %.0f"
          % (w, np.diff(c6[idx]).mean(), (c6 - c4)[idx].mean(), w, np.diff(c8[idx]).mean(), (c8 - c7)[idx].mean()))
print("prerequisite wait of the neighbour loader: median %.0f, mean %.0f, above 500 cycles in %.0f %% of the tiles"
      % (np.median(c5 - c4), (c5 - c4).mean(), 100 * ((c5 - c4) > 500).mean()))
print("tiles finished per tenth of the slot's time:", np.histogram(c3, np.linspace(0, c3.max(), 11))[0])
