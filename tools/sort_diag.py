import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import _lib, pointconv_util as pcu, synth
a, b = synth.frame_pairs(0, 8); a, b = a.cuda(), b.cuda()
L = _lib.lib
for mode in (0, 2, 1):
    L.b200pci_debug_set(17, mode)
    for B in (1, 8):
        keep = pcu._knn(16, a[:B], b[:B], 5, False)
        print(f"sort mode {mode} B={B}: flagged tiles {L.b200pci_debug_get(5)}, flagged queries {L.b200pci_debug_get(6)}")
L.b200pci_debug_set(17, 1)
