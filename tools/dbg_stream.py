import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
from fmcw_radar_processing_b200 import synth
from fmcw_radar_processing_b200.config import fmcw_configurations
from fmcw_radar_processing_b200.parse import make_sxml
from fmcw_radar_processing_b200.streaming import StreamingRecording
NTS, PN, RX = 256, 256, 4
sx = make_sxml(numSamplesPerChirp=NTS, numChirpsPerFrame=PN, numAntennasRx=RX)
cfg = fmcw_configurations(sx); scene = synth.scene_c1(seed=4)
s = StreamingRecording(cfg, synth.default_calib(RX, NTS) / 4095.0)
n=16384; s.reserve(n)
iq = torch.empty((4096, RX, PN, NTS, 2), dtype=torch.int16, device="cuda")
for f0 in range(0, n, 4096):
    tab = synth.scene_tables(scene, cfg["dist_per_bin"], 256, cfg["PRT"], cfg["lambda"], f0, 4096)
    s.h.synth_frames(tab, scene.seed, f0, sigma=scene.sigma, dc=scene.dc, rx_step=scene.rx_step, out=iq)
    s.push_frames(iq, keep_track=False)
buf = torch.empty((2*1024*1024 + 32, 1024), dtype=torch.float32, device="cuda")
L_total, offset, halo = s._layout()
for a, b in s._pieces(2*1024*1024):
    t0=time.perf_counter(); s._load(a, b, halo); torch.cuda.synchronize(); t1=time.perf_counter()
    pm = s.h.stft_local_max(L_total, a); t2=time.perf_counter()
    inf = s.h.info()
    s.h.stft_sharded(L_total, a, pm, buf); torch.cuda.synchronize(); t3=time.perf_counter()
    print(f"piece {a}-{b}: load {1e3*(t1-t0):.2f} ms, local_max {1e3*(t2-t1):.2f} ms (n_refined {inf['n_refined']}), stft {1e3*(t3-t2):.2f} ms, timings {s.h.timings()}")
