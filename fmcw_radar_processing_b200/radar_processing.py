"""Host-side mirror of the reference's function surface (same names, argument meaning, files written and
error behaviour), with the per-frame loop and the STFT block replaced by libfmcw_cuda:

* ``radar_processing(process_animal_activity)``   <- radar-etl-pipeline/radar_processing.m (RP:56)
* ``main(input)``                                 <- radar-etl-pipeline/radar_processing_with_azure.m (RPA:9)

Blob storage I/O (read_data_from_blob_storage / send_json_string_to_blob_storage) is stubbed to local files
as BASELINE.json's north_star asks: inputs are ``radar_data.xml`` + ``radar_data.raw.bin`` in the working
directory (RP:66, RD:15-23), outputs are the JSON files the reference writes into ``pwd``.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import payloads
from .api import FmcwCuda
from .config import array_bin_range, fmcw_configurations
from .parse import f_parse_data2

uploaded = []          # names handed to the upload stub (what the dashboard would fetch)


def read_data_from_blob_storage(workdir="."):
    """Local-file stub of read_data_from_blob_storage.m (RD:3-33): the two files must already be in ``workdir``."""
    for ext in (".xml", ".raw.bin"):
        p = os.path.join(workdir, "radar_data" + ext)
        if not os.path.exists(p):
            raise FileNotFoundError(p)
    return {"xml": "radar_data.xml", "raw": "radar_data.raw.bin"}


def send_json_string_to_blob_storage(filename):
    """Local-file stub of send_json_string_to_blob_storage.m (SJ:4-37): records the upload, never raises
    (the reference prints and swallows upload failures, SJ:34-36)."""
    uploaded.append(os.path.basename(filename))


def send_picture_to_blob_storage(filename):
    """SP:1-80 stubbed to a local file (the PNG stays in the working directory)."""
    return filename


def _speed(cfg, doppler_bin):
    # RP:250 with tgt_doppler_idx = doppler_bin + 1
    return (doppler_bin + 1 - cfg["Doppler_fft_size"] / 2 - 1) * -cfg["fD_per_bin"] * cfg["Hz_to_mps_constant"]


def radar_processing(process_animal_activity, workdir=".", device=0, **cfg_kw):
    """RP:56-612.  ``process_animal_activity`` is compared case-insensitively with 'no' / 'yes' (RP:195, 440);
    any other value runs neither branch.  Returns a dict of what was computed (the reference returns nothing
    and communicates through the files)."""
    fdata = os.path.join(workdir, "radar_data")                      # RP:66
    filename = "radar_data"                                          # RP:68
    frame, frame_count, calib_data, sXML = f_parse_data2(fdata)      # RP:86
    cfg = fmcw_configurations(sXML, **cfg_kw)                        # RP:89-179
    flag = str(process_animal_activity).lower()
    result = {"cfg": cfg, "frame_count": frame_count, "files": []}
    if flag not in ("no", "yes"):
        return result
    h = FmcwCuda(cfg, calib_data, device=device)
    try:
        if flag == "no":
            _branch_no(h, cfg, frame, frame_count, filename, workdir, result)
        else:
            _branch_yes(h, cfg, frame, frame_count, filename, workdir, result)
    finally:
        h.close()
    return result


def _write(workdir, name, fields, result):
    path = os.path.join(workdir, name)
    payloads.write_struct(path, fields)
    send_json_string_to_blob_storage(path)
    result["files"].append(name)


def _branch_no(h, cfg, frame, N, filename, workdir, result):
    """RP:197-436."""
    iq = np.ascontiguousarray(frame)
    out, inten = h.run(iq)                                           # RP:197-299 on the GPU
    info = h.info()
    det = out["detected"].astype(bool)
    L = info["L_total"]
    if L < cfg["window_length"]:
        # RP:276: spectrogram() errors on a signal shorter than the window (the isempty guard at RP:269 is
        # commented out); main() turns that into a failed 'Radar Processing' step (RPA:56-66)
        raise RuntimeError("spectrogram: signal shorter than the window (no target detected in any frame)")
    ncol = info["ncol_local"]
    T, F, nfft, _ = h.stft_axes(L)
    cfg["max_slider_index"] = ncol - cfg["window_length"]            # RP:287
    intensity = inten[:ncol].T                                       # 1024 x ncol view
    _write(workdir, "spectrogram_data.json", payloads.spectrogram_payload(T, F, intensity), result)      # RP:307-328
    # RP:332-348: spectrogram.png = surf(T, F, psd) between ylim [0 150], clim [-40 0], jet -- from the fine-grid band, not
    # from the 671 GB matrix P
    band, _ = h.stft_finegrid(0.0, 150.0, max_rows=2400)
    payloads.write_spectrogram_png(os.path.join(workdir, "spectrogram.png"), band)
    send_picture_to_blob_storage(os.path.join(workdir, "spectrogram.png"))
    result["files"].append("spectrogram.png")
    _write(workdir, filename + "_range_fft_data.json",
           payloads.range_fft_payload(N, array_bin_range(cfg), out["range_max_abs"].T, filename), result)  # RP:355-377
    rng = np.where(det, out["range_bin"] * cfg["dist_per_bin"], 0.0)                                     # RP:248
    spd = np.where(det, _speed(cfg, out["doppler_bin"].astype(np.float64)), 0.0)                         # RP:250
    _write(workdir, filename + "_range_speed_data.json", payloads.range_speed_payload(N, rng, spd, det, filename), result)
    # RP:410-436: range_tx1rx1_complete(:,100) is linear indexing over (chirp, frame)
    PN = cfg["num_chirps_per_frame"]
    if PN * N < 100:
        raise IndexError("Index in position 2 exceeds array bounds")  # what MATLAB raises at RP:411
    spec = h.range_spectrum(iq, 99 // PN, 99 % PN)
    _write(workdir, filename + "_fft_data.json", payloads.fft_payload(spec, filename), result)
    result.update(out=out, intensity=intensity, T=T, frequency=F, info=info, range=rng, speed=spd,
                  strength=np.where(det, out["range_mag"], 0.0))


def _branch_yes(h, cfg, frame, N, filename, workdir, result):
    """RP:440-607: 100-frame batches, at most 4 spectrogram JSONs, NaN for frames without a target."""
    bs, max_plots = cfg["batch_size"], cfg["max_plots"]
    num_batches = math.ceil(N / bs)                                  # RP:190
    strength = np.zeros((1, N)); rng = np.zeros((1, N)); spd = np.zeros((1, N))   # RP:157-159
    plot_counter = 0
    batches = []
    for batch in range(1, num_batches + 1):
        s0, e0 = (batch - 1) * bs + 1, min(batch * bs, N)           # RP:446-447
        iq = np.ascontiguousarray(frame[s0 - 1:e0])
        out = h.process_frames(iq)                                   # RP:457-530
        det = out["detected"].astype(bool)
        sl = slice(s0 - 1, e0)
        strength[0, sl] = np.where(det, out["range_mag"], np.nan)    # RP:500-511, 524-528
        rng[0, sl] = np.where(det, out["range_bin"] * cfg["dist_per_bin"], np.nan)
        spd[0, sl] = np.where(det, _speed(cfg, out["doppler_bin"].astype(np.float64)), np.nan)
        L = int(det.sum()) * cfg["num_chirps_per_frame"]
        if L >= cfg["window_length"]:                                # RP:534
            plot_counter += 1
            if plot_counter <= max_plots:                            # RP:537
                inten = h.stft_frames(e0 - s0 + 1)                   # RP:538-566 on the batch's own signal
                ncol = h.info()["ncol_local"]
                T, F, nfft, _ = h.stft_axes(L)
                name = f"{filename}_spectrogram_batch_{batch}.json"  # RP:587
                _write(workdir, name, payloads.batch_spectrogram_payload(T, F, inten[:ncol].T, batch, s0, e0, filename), result)
                batches.append(dict(batch=batch, start_frame=s0, end_frame=e0, T=T, frequency=F, intensity=inten[:ncol].T))
            else:
                break                                                # RP:599
    result.update(batches=batches, strength=strength, range=rng, speed=spd)


def main(input):
    """radar_processing_with_azure.m ``main(input)`` (RPA:9-100): same step list, same status/message strings,
    errors are reported in the result, never raised.  ``input`` is a dict; optional keys ``workdir`` / ``device``
    are extensions for local use."""
    steps = []
    flag = "no"                                                      # RPA:15
    if isinstance(input, dict) and "processAnimalActivity" in input:
        flag = input["processAnimalActivity"]                        # RPA:16-18
    workdir = input.get("workdir", ".") if isinstance(input, dict) else "."
    device = input.get("device", 0) if isinstance(input, dict) else 0
    try:                                                             # RPA:25-45
        read_data_from_blob_storage(workdir)
        steps.append({"step": "Read Files", "status": "success",
                      "message": "Files downloaded from Azure Blob Storage successfully."})
    except Exception as ME:
        return {"status": "error", "message": "Failed at reading files from blob storage.",
                "steps": [{"step": "Read Files", "status": "error", "message": str(ME)}]}
    try:                                                             # RPA:48-66
        radar_processing(flag, workdir=workdir, device=device)
        steps.append({"step": "Radar Processing", "status": "success", "message": "Radar data processed successfully."})
    except Exception as ME:
        steps.append({"step": "Radar Processing", "status": "error", "message": str(ME)})
        return {"status": "error", "message": "Failed at radar processing step.", "steps": steps}
    # RPA:67 is a lost comment marker that makes the MATLAB original throw here; the intended flow continues
    steps.append({"step": "Upload JSON", "status": "success",
                  "message": "Processed JSON uploaded to Azure Blob Storage."})      # RPA:68-74
    return {"status": "success", "message": "All steps completed successfully.", "steps": steps}   # RPA:95-99
