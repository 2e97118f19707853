import sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import torch
from fmcw_radar_processing_b200 import synth, _lib
from fmcw_radar_processing_b200.api import FmcwCuda, _ptr
from fmcw_radar_processing_b200.config import fmcw_configurations
from fmcw_radar_processing_b200.parse import make_sxml
dev = torch.device("cuda", 0)
NTS, PN, F = 128, 64, 500
sx = make_sxml(numSamplesPerChirp=NTS, numChirpsPerFrame=PN, numAntennasRx=1)
calib = synth.default_calib(1, NTS) / 4095.0
for win, ov in ((20, 19), (32, 16)):
    cfg = fmcw_configurations(sx, window_length=win, overlap=ov)
    NH = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    hs = [FmcwCuda(cfg, calib, device=0, torch_stream_sync=False) for _ in range(NH)]
    if len(sys.argv) > 2 and sys.argv[2] == 'graph':
        for h in hs: h.set_option(_lib.OPT_RUN_GRAPH, 1)
    sc = synth.scene_c1(seed=1000)
    tab = synth.scene_tables(sc, cfg["dist_per_bin"], 256, cfg["PRT"], cfg["lambda"], 0, F)
    iq = torch.empty((F, 1, PN, NTS, 2), dtype=torch.int16, device=dev)
    hs[0].synth_frames(tab, sc.seed, 0, sigma=sc.sigma, dc=sc.dc, rx_step=sc.rx_step, out=iq)
    torch.cuda.synchronize()
    outs = [h.alloc_frame_out(F, device=dev) for h in hs]
    intens = [torch.empty((max(1, h.max_cols(F)), h.nq), dtype=torch.float32, device=dev) for h in hs]
    for _ in range(4):
        for h, o, it in zip(hs, outs, intens): h.run(iq, o, it)
    torch.cuda.synchronize()
    N = 512
    t0 = time.perf_counter()
    for i in range(N):
        k = i % NH
        hs[k].run(iq, outs[k], intens[k])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"win {win}: python h.run enqueue {1e6*(t1-t0)/N:.1f} us per radar, with final sync {1e6*(t2-t0)/N:.1f} us per radar")
    # raw ctypes calls with prebuilt structs
    fos = [h._frame_struct(o) for h, o in zip(hs, outs)]
    sos = [h._stft_struct(it, 0) for h, it in zip(hs, intens)]
    p = _ptr(iq)
    t0 = time.perf_counter()
    for i in range(N):
        k = i % NH
        hs[k].lib.fmcw_run(hs[k]._h, p, F, C.byref(fos[k]), C.byref(sos[k]))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"win {win}: raw fmcw_run enqueue {1e6*(t1-t0)/N:.1f} us per radar, with final sync {1e6*(t2-t0)/N:.1f} us per radar")
    t0 = time.perf_counter()
    for i in range(N):
        k = i % NH
        hs[k].lib.fmcw_run(hs[k]._h, p, F, C.byref(fos[k]), C.byref(sos[k]))
        if i >= NH: hs[k].info()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"win {win}: raw fmcw_run + info per radar {1e6*(t1-t0)/N:.1f} us")
    for h in hs: h.close()
