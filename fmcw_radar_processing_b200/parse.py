"""Recording I/O: the local-file stand-in for the reference's unshipped vendor parser
``f_parse_data2`` and File Exchange ``xml2struct`` (call sites RP:81-86).

The reference reads ``radar_data.xml`` + ``radar_data.raw.bin`` from ``pwd`` (RP:66, RD:15-23).  The vendor
``.raw.bin`` layout is not documented in the reference, so this module DEFINES a container for the same
content (parity unpinned):

    magic  b"FMCWRAW1"
    uint32 n_frames, n_rx, chirps_per_frame, samples_per_chirp, n_cal (calibration samples per RX)
    float64 calib[2 * n_rx * n_cal]       row vector [I_rx1 Q_rx1 I_rx2 Q_rx2 ...] in ADC codes (RP:167-172)
    int16  iq[n_frames][n_rx][chirps][samples][2]   12-bit ADC codes, I/Q interleaved

``f_parse_data2`` returns the int16 block untouched (it is the library's input format); the division by
``adc_scale`` = 4095 that turns codes into the reference's normalised doubles happens inside the kernels.
"""
from __future__ import annotations

import struct
import xml.etree.ElementTree as ET

import numpy as np

MAGIC = b"FMCWRAW1"


def make_sxml(chirpDuration_ns=300000, lowerFrequency_kHz=24025000, upperFrequency_kHz=24225000, numAntennasTx=1,
              numAntennasRx=1, numSamplesPerChirp=128, numChirpsPerFrame=64, samplerateHz=426666):
    """xml2struct-shaped dict holding the ``.Text`` leaves RP:94-115 reads (dashboard values)."""
    t = lambda v: {"Text": str(v)}
    return {"Device": {
        "BaseEndpoint": {"chirpDuration_ns": t(chirpDuration_ns),
                         "DeviceInfo": {"numAntennasTx": t(numAntennasTx), "numAntennasRx": t(numAntennasRx)},
                         "FrameFormat": {"numSamplesPerChirp": t(numSamplesPerChirp),
                                         "numChirpsPerFrame": t(numChirpsPerFrame)}},
        "FmcwEndpoint": {"FmcwConfiguration": {"upperFrequency_kHz": t(upperFrequency_kHz),
                                               "lowerFrequency_kHz": t(lowerFrequency_kHz)}},
        "AdcxmcEndpoint": {"AdcxmcConfiguration": {"samplerateHz": t(samplerateHz)}}}}


def _to_xml(name, node):
    el = ET.Element(name)
    for k, v in node.items():
        if k == "Text":
            el.text = v
        else:
            el.append(_to_xml(k, v))
    return el


def xml2struct(path):
    """Nested dict with ``Text`` leaves, like File Exchange #28518 (RP:73-83)."""
    def conv(el):
        d = {c.tag: conv(c) for c in el}
        if el.text is not None and el.text.strip():
            d["Text"] = el.text.strip()
        return d
    root = ET.parse(path).getroot()
    return {root.tag: conv(root)}


def write_recording(fdata, iq, calib_codes, sxml):
    """Writes ``<fdata>.xml`` and ``<fdata>.raw.bin``."""
    iq = np.ascontiguousarray(iq, dtype=np.int16)
    n_frames, n_rx, PN, NTS, two = iq.shape
    assert two == 2
    calib = np.ascontiguousarray(calib_codes, dtype=np.float64).reshape(-1)
    n_cal = calib.size // (2 * n_rx)
    ET.ElementTree(_to_xml("Device", sxml["Device"])).write(fdata + ".xml")
    with open(fdata + ".raw.bin", "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<5I", n_frames, n_rx, PN, NTS, n_cal))
        f.write(calib.tobytes())
        f.write(iq.tobytes())


def f_parse_data2(fdata, adc_scale=4095.0):
    """``[frame, frame_count, calib_data, sXML] = f_parse_data2(fdata)`` (RP:86).  ``frame`` is the int16
    block [frame][rx][chirp][sample][2] (memory-mapped); ``calib_data`` is normalised like the reference's."""
    sxml = xml2struct(fdata + ".xml")
    with open(fdata + ".raw.bin", "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{fdata}.raw.bin: not an FMCWRAW1 container")
        n_frames, n_rx, PN, NTS, n_cal = struct.unpack("<5I", f.read(20))
        calib = np.frombuffer(f.read(8 * 2 * n_rx * n_cal), dtype=np.float64) / adc_scale
        off = f.tell()
    frame = np.memmap(fdata + ".raw.bin", dtype=np.int16, mode="r", offset=off, shape=(n_frames, n_rx, PN, NTS, 2))
    return frame, n_frames, calib, sxml
