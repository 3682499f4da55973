import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import epi_probe as E
for M in (32768, 37888):
    for nb in (1, 6):
        us, tf = E.run(M, 512, 512, 0, nbuf=nb)
        print(f"fwd 512x512 M={M} nbuf={nb} full: {us:.1f} us {tf:.0f} TF/s")
        us, tf = E.run(M, 512, 512, 0, nbuf=nb, store=False, mask=False)
        print(f"fwd 512x512 M={M} nbuf={nb} nostore: {us:.1f} us {tf:.0f} TF/s")
        us, tf = E.run(M, 512, 512, 1, nbuf=nb)
        print(f"dgrad 512x512 M={M} nbuf={nb} full: {us:.1f} us {tf:.0f} TF/s")
