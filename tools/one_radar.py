import sys
sys.path.insert(0, "/root/repo")
import torch
from fmcw_radar_processing_b200 import synth
from fmcw_radar_processing_b200.api import FmcwCuda
from fmcw_radar_processing_b200.config import fmcw_configurations
from fmcw_radar_processing_b200.parse import make_sxml
win, ov = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20, 19)
dev = torch.device("cuda", 0)
NTS, PN, F = 128, 64, 500
sx = make_sxml(numSamplesPerChirp=NTS, numChirpsPerFrame=PN, numAntennasRx=1)
calib = synth.default_calib(1, NTS) / 4095.0
cfg = fmcw_configurations(sx, window_length=win, overlap=ov)
h = FmcwCuda(cfg, calib, device=0)
sc = synth.scene_c1(seed=1000)
tab = synth.scene_tables(sc, cfg["dist_per_bin"], 256, cfg["PRT"], cfg["lambda"], 0, F)
iq = torch.empty((F, 1, PN, NTS, 2), dtype=torch.int16, device=dev)
h.synth_frames(tab, sc.seed, 0, sigma=sc.sigma, dc=sc.dc, rx_step=sc.rx_step, out=iq)
out = h.alloc_frame_out(F, device=dev)
inten = torch.empty((max(1, h.max_cols(F)), h.nq), dtype=torch.float32, device=dev)
for _ in range(3):
    h.run(iq, out, inten)
    torch.cuda.synchronize()
print(h.info(), h.timings())
