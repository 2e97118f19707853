"""GPU tests of the reference-facing surface: radar_processing('no'/'yes') and main() through the JSON
payloads the dashboard consumes, and the sharded STFT path (emulated ranks on one GPU)."""
import json
import os

import numpy as np
import pytest

from fmcw_radar_processing_b200 import synth
from fmcw_radar_processing_b200.parse import write_recording
from oracle import fmcw_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _recording(tmp_path, case):
    write_recording(str(tmp_path / "radar_data"), case["iq"], case["calib_codes"], case["sxml"])


def test_no_branch_payloads_match_oracle(tmp_path):
    from fmcw_radar_processing_b200.radar_processing import main, uploaded
    case = H.make_case(n_frames=30, NTS=128, PN=64)
    case["iq"][[4, 17]] = 2048                     # two frames without a target
    _recording(tmp_path, case)
    ref = H.oracle_no(case)
    del uploaded[:]
    r = main({"processAnimalActivity": "No", "workdir": str(tmp_path)})        # strcmpi (RP:195)
    assert r["status"] == "success", r
    assert [s["step"] for s in r["steps"]] == ["Read Files", "Radar Processing", "Upload JSON"]
    assert uploaded == ["spectrogram_data.json", "radar_data_range_fft_data.json", "radar_data_range_speed_data.json",
                        "radar_data_fft_data.json"]
    sp = json.load(open(tmp_path / "spectrogram_data.json"))
    assert list(sp) == ["time", "frequency", "intensity", "title", "xLabel", "yLabel"]
    st = ref["stft"]
    assert np.allclose(sp["time"], st["T"], rtol=1e-13) and np.allclose(sp["frequency"], st["frequency"], rtol=1e-13)
    inten = np.array(sp["intensity"], dtype=np.float64)
    assert inten.shape == st["intensity"].shape == (1024, 28 * 64 - 19)
    e_db, _ = H.spectrogram_errors(inten, st["intensity"])
    assert e_db < 1e-3
    rf = json.load(open(tmp_path / "radar_data_range_fft_data.json"))
    assert np.allclose(rf["time_axis"], np.arange(30) * 0.15) and np.allclose(rf["array_bin_range"], case["ocfg"].array_bin_range)
    e_db, _ = H.db_errors(np.array(rf["range_tx1rx1_max_abs"]), ref["range_tx1rx1_max_abs"])
    assert e_db < 1e-3 and rf["filename"] == "radar_data"
    rs = json.load(open(tmp_path / "radar_data_range_speed_data.json"))
    rng = np.array(rs["range"])
    assert rng.shape == (30, 30)                   # lastDetectedFrame x frame_count (RP:245-250 quirk)
    # identical bins -> identical metres / m/s up to jsonencode's 15 significant digits
    assert np.allclose(rng[:, 0], ref["range"], rtol=1e-14, atol=0) and np.count_nonzero(rng[:, 1:]) == 0
    assert np.allclose(np.array(rs["speed"])[:, 0], ref["speed"], rtol=1e-14, atol=0)
    ff = json.load(open(tmp_path / "radar_data_fft_data.json"))
    assert ff["frame_index"] == 100 and ff["range_bins"] == list(range(256)) and len(ff["magnitude"]) == 256
    # RP:332-348: spectrogram.png from the fine-grid psd band (0-150 Hz, clim [-40 0], jet)
    import struct
    import zlib
    png = open(tmp_path / "spectrogram.png", "rb").read()
    assert png[:8] == b"\x89PNG\r\n\x1a\n" and png[12:16] == b"IHDR"
    w, hgt = struct.unpack(">II", png[16:24])
    assert w == 28 * 64 - 19 and 100 < hgt <= 2400
    ilen = struct.unpack(">I", png[33:37])[0]
    raw = zlib.decompress(png[41:41 + ilen])
    assert len(raw) == hgt * (1 + 3 * w)
    px = np.frombuffer(raw, dtype=np.uint8).reshape(hgt, 1 + 3 * w)[:, 1:].reshape(hgt, w, 3)
    assert (px[-1, :, 0] > 100).mean() > 0.9       # the bottom row (lowest frequencies) is the hot end of jet: red


def test_no_branch_without_any_target_is_a_failed_step(tmp_path):
    from fmcw_radar_processing_b200.radar_processing import main
    case = H.make_case(n_frames=3, NTS=64, PN=16)
    case["iq"][:] = 2048
    _recording(tmp_path, case)
    r = main({"processAnimalActivity": "no", "workdir": str(tmp_path)})
    assert r["status"] == "error" and r["message"] == "Failed at radar processing step."      # RP:276 errors, RPA:56-66


def test_other_flag_runs_neither_branch(tmp_path):
    from fmcw_radar_processing_b200.radar_processing import radar_processing
    case = H.make_case(n_frames=2, NTS=64, PN=16)
    _recording(tmp_path, case)
    assert radar_processing("maybe", workdir=str(tmp_path))["files"] == []


def test_yes_branch_batches_match_oracle(tmp_path):
    from fmcw_radar_processing_b200.radar_processing import radar_processing
    case = H.make_case(n_frames=520, NTS=64, PN=16)
    case["iq"][130:180] = 2048                     # a gap inside batch 2
    _recording(tmp_path, case)
    frames, n, calib, sx = O.f_parse_data2(case["iq"], case["calib_codes"], case["sxml"])
    ref = O.radar_processing_yes(frames, calib, sx)
    r = radar_processing("YES", workdir=str(tmp_path))
    assert r["files"] == [f"radar_data_spectrogram_batch_{b}.json" for b in (1, 2, 3, 4)]     # RP:443, 587
    for got, exp in zip(r["batches"], ref["batches"]):
        assert (got["batch"], got["start_frame"], got["end_frame"]) == (exp["batch"], exp["start_frame"], exp["end_frame"])
        assert got["intensity"].shape == exp["intensity"].shape
        e_db, _ = H.spectrogram_errors(got["intensity"], exp["intensity"])
        assert e_db < 1e-3
        assert np.allclose(got["T"], exp["T"], rtol=1e-13)
    assert np.array_equal(np.isnan(r["range"]), np.isnan(ref["range"]))       # NaN fill (RP:524-528); frames > 500 stay 0 (RP:599)
    assert np.array_equal(np.nan_to_num(r["range"]), np.nan_to_num(ref["range"]))
    assert np.array_equal(np.nan_to_num(r["speed"]), np.nan_to_num(ref["speed"]))
    b2 = json.load(open(tmp_path / "radar_data_spectrogram_batch_2.json"))
    assert b2["title"] == "Spectrogram - Batch 2" and b2["start_frame"] == 101 and b2["end_frame"] == 200


@pytest.mark.parametrize("split", [(20, 20), (37, 3), (1, 39), (13, 13, 14)])
def test_sharded_stft_equals_single_gpu(split):
    """The C-ABI sharded path (counts -> halo -> local max -> global max -> sharded STFT) with the ranks
    emulated as handles on one GPU reproduces the single-handle result column for column."""
    from fmcw_radar_processing_b200.api import FmcwCuda
    n = sum(split)
    case = H.make_case(n_frames=n, NTS=128, PN=64)
    case["iq"][5] = 2048
    h0 = FmcwCuda(case["cfg"], case["calib"])
    out0, inten0 = h0.run(case["iq"])
    info0 = h0.info()
    win = case["cfg"]["window_length"]
    hs, Ls, heads = [], [], []
    f0 = 0
    for k in split:
        h = FmcwCuda(case["cfg"], case["calib"])
        h.process_frames(np.ascontiguousarray(case["iq"][f0:f0 + k]))
        L = h.info()["L_local"]
        head = np.zeros(min(L, win - 1), dtype=np.float64)
        if head.size:
            h.get_slow_time(head, 0, head.size)
        hs.append(h); Ls.append(L); heads.append(head)
        f0 += k
    L_total = sum(Ls)
    assert L_total == info0["L_total"]
    offs = np.concatenate([[0], np.cumsum(Ls)[:-1]]).astype(int)
    maxes = []
    for r, h in enumerate(hs):
        halo = np.concatenate(heads[r + 1:] + [np.zeros(0, np.float64)])[:win - 1].astype(np.float64)
        h.set_halo(np.ascontiguousarray(halo), halo.size)
        maxes.append(h.stft_local_max(L_total, int(offs[r])))
    pmax = max(maxes)
    assert pmax == pytest.approx(info0["pmax_raw"], rel=1e-6)
    cols = []
    for r, h in enumerate(hs):
        buf = np.empty((max(1, Ls[r]), 1024), dtype=np.float32)
        h.stft_sharded(L_total, int(offs[r]), info0["pmax_raw"], buf)
        inf = h.info()
        assert inf["col_begin"] == sum(c.shape[0] for c in cols)
        cols.append(buf[:inf["ncol_local"]])
        h.close()
    got = np.concatenate(cols)
    assert got.shape[0] == info0["ncol_total"]
    assert np.array_equal(got, inten0[:info0["ncol_total"]])
    h0.close()


@pytest.mark.parametrize("split", [(20, 20), (37, 3), (13, 0, 14, 13)])
def test_async_sharded_path_equals_single_gpu(split):
    """fmcw_shard_pack / fmcw_shard_plan / fmcw_shard_stft: the device-side hand-offs (headers, halo, offsets,
    global max) with the all-gather and all-reduce emulated by torch ops on one GPU."""
    import torch
    from fmcw_radar_processing_b200.api import FmcwCuda
    n = sum(split)
    case = H.make_case(n_frames=n, NTS=128, PN=64)
    case["iq"][5] = 2048
    iq_d = torch.from_numpy(case["iq"]).cuda()
    h0 = FmcwCuda(case["cfg"], case["calib"])
    out0, inten0 = h0.run(iq_d)
    info0 = h0.info()
    win = case["cfg"]["window_length"]
    world = len(split)
    hs, msgs, outs = [], [], []
    f0 = 0
    for k in split:
        h = FmcwCuda(case["cfg"], case["calib"])
        if k:
            outs.append(h.process_frames(iq_d[f0:f0 + k].contiguous()))
        else:
            outs.append(h.process_frames(iq_d[:0].contiguous()))
        msg = torch.zeros(1 + win - 1, dtype=torch.float64, device="cuda")
        h.shard_pack(msg)
        h.synchronize()
        hs.append(h); msgs.append(msg)
        f0 += k
    gathered = torch.cat(msgs).contiguous()
    maxes = []
    for r, h in enumerate(hs):
        m = torch.zeros(1, dtype=torch.float64, device="cuda")
        h.shard_plan(gathered, world, r, m)
        h.synchronize()
        maxes.append(m)
    gmax = torch.stack(maxes).max().reshape(1).contiguous()
    assert float(gmax) == pytest.approx(info0["pmax_raw"], rel=1e-6)
    gmax[0] = info0["pmax_raw"]
    cols = []
    for r, h in enumerate(hs):
        buf = torch.empty((max(1, split[r] * 64), 1024), dtype=torch.float32, device="cuda")
        h.shard_stft(gmax, buf)
        inf = h.info()
        assert inf["L_total"] == info0["L_total"] and inf["col_begin"] == sum(c.shape[0] for c in cols)
        cols.append(buf[:inf["ncol_local"]].cpu().numpy())
        h.close()
    got = np.concatenate(cols)
    assert got.shape[0] == info0["ncol_total"]
    assert np.array_equal(got, inten0[:info0["ncol_total"]].cpu().numpy())
    h0.close()


@pytest.mark.parametrize("split", [(20, 20), (37, 3), (13, 0, 14, 13), (40,)])
def test_mailbox_sharded_path_equals_single_gpu(split):
    """fmcw_mailbox_post_heads / fmcw_mailbox_plan / fmcw_mailbox_stft: the peer-memory hand-offs (headers, halo,
    offsets, global max exchanged by stores + flags inside the kernels), with the ranks emulated as handles on one
    GPU whose mailboxes are ordinary device buffers.  Four passes in a row reuse the mailboxes (step numbers): the first
    two launch the exchange kernels one by one, the third captures them as one CUDA graph, the fourth replays it."""
    import torch
    from fmcw_radar_processing_b200.api import FmcwCuda
    n = sum(split)
    case = H.make_case(n_frames=n, NTS=128, PN=64)
    case["iq"][5] = 2048
    iq_d = torch.from_numpy(case["iq"]).cuda()
    h0 = FmcwCuda(case["cfg"], case["calib"])
    out0, inten0 = h0.run(iq_d)
    info0 = h0.info()
    ref = inten0[:info0["ncol_total"]].cpu().numpy()
    world = len(split)
    hs = [FmcwCuda(case["cfg"], case["calib"]) for _ in split]
    boxes = [torch.zeros((hs[0].mailbox_bytes() + 7) // 8, dtype=torch.float64, device="cuda") for _ in split]
    torch.cuda.synchronize()
    ptrs = [b.data_ptr() for b in boxes]
    for step in (1, 2, 3, 4):
        f0 = 0
        for r, (h, k) in enumerate(zip(hs, split)):
            h.process_frames(iq_d[f0:f0 + k].contiguous())
            h.mailbox_post_heads(ptrs, r, step)
            f0 += k
        for r, h in enumerate(hs):
            h.mailbox_plan(ptrs, r, step)
        bufs = []
        for r, h in enumerate(hs):
            buf = torch.empty((max(1, split[r] * 64), 1024), dtype=torch.float32, device="cuda")
            h.mailbox_stft(ptrs, r, step, buf)
            bufs.append(buf)
        cols = []
        for r, h in enumerate(hs):
            inf = h.info()
            assert inf["L_total"] == info0["L_total"] and inf["col_begin"] == sum(c.shape[0] for c in cols)
            assert inf["pmax_raw"] == pytest.approx(info0["pmax_raw"], rel=1e-6)
            cols.append(bufs[r][:inf["ncol_local"]].cpu().numpy())
        got = np.concatenate(cols)
        assert got.shape == ref.shape
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), fin)
        assert np.max(np.abs(got[fin] - ref[fin])) <= 2e-4       # the global max agrees to 1e-6 relative = 9e-6 dB
    for h in hs:
        h.close()
    h0.close()


def test_mailbox_argument_errors():
    """The mailbox entry points validate world / rank / step / pointers before anything is queued."""
    import torch
    from fmcw_radar_processing_b200 import FmcwError
    from fmcw_radar_processing_b200.api import FmcwCuda
    case = H.make_case(n_frames=4, NTS=128, PN=64)
    h = FmcwCuda(case["cfg"], case["calib"])
    box = torch.zeros((h.mailbox_bytes() + 7) // 8, dtype=torch.float64, device="cuda")
    ptrs = [box.data_ptr()]
    with pytest.raises(FmcwError) as ei:                    # no frames processed yet
        h.mailbox_post_heads(ptrs, 0, 1)
    assert ei.value.status == 9
    h.process_frames(torch.from_numpy(case["iq"]).cuda())
    with pytest.raises(FmcwError) as ei:                    # rank >= world
        h.mailbox_post_heads(ptrs, 1, 1)
    assert ei.value.status == 7
    with pytest.raises(FmcwError) as ei:                    # step numbers start at 1
        h.mailbox_post_heads(ptrs, 0, 0)
    assert ei.value.status == 7
    with pytest.raises(FmcwError) as ei:                    # null mailbox
        h.mailbox_post_heads([0], 0, 1)
    assert ei.value.status == 2
    with pytest.raises(FmcwError) as ei:                    # STFT before the plan
        h.mailbox_stft(ptrs, 0, 1, torch.empty((300, 1024), dtype=torch.float32, device="cuda"))
    assert ei.value.status == 9
    assert h.mailbox_bytes() >= 64 * 24 + 64 * 8 * case["cfg"]["window_length"]
    h.close()


def test_fleet_of_radars_equals_one_by_one():
    """C5: independent radars through a pool of handles on concurrent streams give, radar by radar, exactly what a
    single handle gives (own nfft, own maximum; recordings of different lengths, one without any target)."""
    import torch
    from fmcw_radar_processing_b200.api import FmcwCuda
    from fmcw_radar_processing_b200.fleet import Fleet
    cases = [H.make_case(n_frames=n, NTS=128, PN=64, seed=100 + i) for i, n in enumerate((12, 30, 7, 21, 16, 9))]
    cases[2]["iq"][:] = 2048                                   # a radar that never detects anything
    cfg, calib = cases[0]["cfg"], cases[0]["calib"]
    recs = [torch.from_numpy(c["iq"]).cuda() for c in cases]
    fleet = Fleet(cfg, calib, n_handles=3)
    h = FmcwCuda(cfg, calib)
    want = []
    for rec in recs:
        out, inten = h.run(rec)
        info = h.info()
        want.append(({k: v.clone() for k, v in out.items()}, inten[:info["ncol_local"]].clone(), info))
    # passes 1-2 run launch by launch, pass 3 records every recording's run as one CUDA graph, passes 4-5 replay the graphs
    for p in range(5):
        for g in fleet._bufs.values():                         # results of the previous pass must not survive by accident
            g[2].fill_(-1.0)
        got = fleet.run(recs)
        for i, (out, inten, info) in enumerate(want):
            g = got[i]
            assert g["info"]["n_detected"] == info["n_detected"] and g["info"]["nfft"] == info["nfft"], (p, i)
            assert g["info"]["pmax_raw"] == info["pmax_raw"] and g["ncol"] == info["ncol_local"], (p, i)
            assert g["info"]["n_frames"] == info["n_frames"] and g["info"]["L_local"] == info["L_local"]
            for k in ("detected", "range_bin", "doppler_bin", "range_mag", "range_max_abs"):
                assert torch.equal(g[k], out[k]), (p, i, k)
            n = info["ncol_local"]
            assert torch.equal(torch.nan_to_num(g["intensity"][:n], nan=7.0), torch.nan_to_num(inten, nan=7.0)), (p, i)
        assert got[2]["info"]["n_detected"] == 0 and got[2]["ncol"] == 0
    h.close()
    fleet.close()


def test_run_graph_option_replays_and_survives_buffer_growth():
    """FMCW_OPT_RUN_GRAPH on one handle: the third run on a buffer set is recorded, later ones replay it; a bigger recording in
    between moves the scratch buffers, which drops the recorded graphs instead of replaying stale addresses."""
    import torch
    from fmcw_radar_processing_b200 import _lib
    from fmcw_radar_processing_b200.api import FmcwCuda
    small = H.make_case(n_frames=24, NTS=128, PN=64, seed=5)
    big = H.make_case(n_frames=90, NTS=128, PN=64, seed=6)
    cfg, calib = small["cfg"], small["calib"]
    ref = FmcwCuda(cfg, calib)
    iq_s, iq_b = torch.from_numpy(small["iq"]).cuda(), torch.from_numpy(big["iq"]).cuda()
    out_s, inten_s = ref.run(iq_s)
    n_s = ref.info()["ncol_local"]
    want_s = inten_s[:n_s].clone()
    out_b, inten_b = ref.run(iq_b)
    n_b = ref.info()["ncol_local"]
    want_b = inten_b[:n_b].clone()
    h = FmcwCuda(cfg, calib)
    h.set_option(_lib.OPT_RUN_GRAPH, 1)
    o1 = h.alloc_frame_out(24, device=iq_s.device)
    i1 = torch.empty((max(1, h.max_cols(24)), h.nq), dtype=torch.float32, device=iq_s.device)
    o2 = h.alloc_frame_out(90, device=iq_s.device)
    i2 = torch.empty((max(1, h.max_cols(90)), h.nq), dtype=torch.float32, device=iq_s.device)
    for p in range(5):                                         # plain, plain, recorded, replayed, replayed
        i1.fill_(-1.0)
        h.run(iq_s, o1, i1)
        assert h.info()["ncol_local"] == n_s
        assert torch.equal(i1[:n_s], want_s), p
        assert torch.equal(o1["range_bin"], out_s["range_bin"])
    h.run(iq_b, o2, i2)                                        # grows xc / o_slow64 / ...: the small recording's graph is stale now
    assert torch.equal(i2[:n_b], want_b)
    for p in range(4):
        i1.fill_(-1.0)
        h.run(iq_s, o1, i1)
        assert torch.equal(i1[:n_s], want_s), p
        i2.fill_(-1.0)
        h.run(iq_b, o2, i2)
        assert h.info()["ncol_local"] == n_b
        assert torch.equal(i2[:n_b], want_b), p
    t = h.timings()                                            # replayed runs carry no stage events
    assert t["chain_ms"] == 0.0 and t["stft_main_ms"] == 0.0
    h.close()
    ref.close()


def test_all_rx_streams_match_the_oracle_per_antenna():
    """rx_select = all: one stream per antenna on the same frame buffer; antenna r equals the oracle run on
    that antenna, from host buffers and from a device tensor."""
    import torch
    from fmcw_radar_processing_b200.fleet import AllRx
    case = H.make_case(n_frames=16, NTS=128, PN=64, n_rx=3)
    a = AllRx(case["cfg"], case["calib"])
    res_h = a.run(case["iq"])
    res_d = a.run(torch.from_numpy(case["iq"]).cuda())
    for r in range(3):
        ref = H.oracle_no(case, rx_select=r + 1)
        for res in (res_h, res_d):
            out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in res[r].items() if k not in ("info",)}
            d = ref["detected"]
            assert np.array_equal(out["detected"].astype(bool), d) and np.array_equal(out["range_bin"][d], ref["range_idx"][d] - 1)
            nc = res[r]["ncol"]
            H.assert_spectrogram_contract(out["intensity"][:nc].T, ref["stft"]["intensity"])
    assert not np.array_equal(res_h[0]["intensity"][:100], res_h[1]["intensity"][:100])      # the antennas differ (phase offset)
    a.close()


def test_streaming_recording_equals_one_shot():
    """streaming.py (BASELINE configs[3] driver): frames pushed in chunks, the STFT in pieces through a small reusable
    buffer -- the same result as one fmcw_run over the whole recording (spectrogram, track, nfft, max)."""
    import torch
    from fmcw_radar_processing_b200.api import FmcwCuda
    from fmcw_radar_processing_b200.streaming import StreamingRecording
    case = H.make_case(n_frames=50, NTS=128, PN=64)
    case["iq"][[7, 31]] = 2048
    iq_d = torch.from_numpy(case["iq"]).cuda()
    h0 = FmcwCuda(case["cfg"], case["calib"])
    out0, inten0 = h0.run(iq_d)
    info0 = h0.info()
    ref = inten0[:info0["ncol_total"]].cpu().numpy()
    s = StreamingRecording(case["cfg"], case["calib"])
    s.reserve(50)
    for f0 in range(0, 50, 12):
        s.push_frames(iq_d[f0:f0 + 12].contiguous())
    got = np.full_like(ref, np.nan)
    buf = torch.empty((700 + 20, 1024), dtype=torch.float32, device="cuda")

    def consumer(c0, n, b):
        got[c0:c0 + n] = b[:n].cpu().numpy()

    r = s.stft(buf, piece_cols=700, consumer=consumer)
    assert r["L_total"] == info0["L_total"] and r["ncol_total"] == info0["ncol_total"] == r["ncol_local"] and r["pieces"] == 5
    assert r["pmax_raw"] == pytest.approx(info0["pmax_raw"], rel=1e-6)
    assert np.array_equal(np.isfinite(got), np.isfinite(ref)) and np.nanmax(np.abs(got - ref)) <= 2e-4
    rb = np.concatenate([t["range_bin"] for t in s.track])
    assert np.array_equal(rb, out0["range_bin"].cpu().numpy())
    s.close(); h0.close()
