"""Output side of `inference.py:208-224`: 10-s truncation and 16-bit PCM WAV files.

The reference writes each int16 clip with `soundfile.write(path, wav, samplerate=sr)`; for an int16 array libsndfile
emits the canonical 44-byte RIFF/WAVE header (PCM format tag 1, one `fmt ` chunk of 16 bytes, one `data` chunk)
followed by the little-endian samples.  soundfile is not in this image, so the writer below produces those bytes
directly (tests/test_host_cpu.py checks them against a hand-built header and Python's `wave` module).
"""
import os
import struct

import numpy as np


def wav_bytes_pcm16(samples, sr=16000):
    """int16 array [T] (mono) or [T, C] -> bytes of a PCM_16 WAV file."""
    a = np.asarray(samples)
    if a.dtype != np.int16:
        raise TypeError("PCM_16 writer takes int16 samples (vocoder_infer already returns int16), got %s" % a.dtype)
    if a.ndim == 1:
        ch = 1
    elif a.ndim == 2:
        ch = a.shape[1]
    else:
        raise ValueError("samples must be [T] or [T, channels]")
    data = np.ascontiguousarray(a).astype("<i2", copy=False).tobytes()
    block = 2 * ch
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE"
    hdr += b"fmt " + struct.pack("<IHHIIHH", 16, 1, ch, int(sr), int(sr) * block, block, 16)
    hdr += b"data" + struct.pack("<I", len(data))
    return hdr + data


def write_wav_pcm16(path, samples, sr=16000):
    with open(path, "wb") as f:
        f.write(wav_bytes_pcm16(samples, sr))
    return path


def save_clips(waves, output_dir, sr=16000, seconds=10.0, prefix="output_"):
    """inference.py:208,221-222: every clip truncated to `seconds` and written as {output_dir}/output_{j}.wav."""
    os.makedirs(output_dir, exist_ok=True)
    n = int(sr * seconds)
    paths = []
    for j, w in enumerate(waves):
        paths.append(write_wav_pcm16(os.path.join(output_dir, "%s%d.wav" % (prefix, j)), np.asarray(w)[:n], sr))
    return paths
