"""In-memory stand-in for ``digital_rf.DigitalRFReader`` (not installed here): just the four methods
``DrfInput`` calls (drfProc.py:63-126)."""
import numpy as np


class FakeReader:
    def __init__(self, channels, sample_rate=1000000, first_sample=0, int16=False):
        self.channels = {k: np.asarray(v) for k, v in channels.items()}
        self.sr = int(sample_rate)
        self.first = int(first_sample)
        self.int16 = int16
        self.reads = []

    def get_channels(self):
        return list(self.channels)

    def get_properties(self, chan):
        d = self.channels[chan]
        props = {"sample_rate_numerator": self.sr, "sample_rate_denominator": 1,
                 "num_subchannels": 1 if d.ndim == 1 else d.shape[1]}
        if self.int16:
            props.update(H5Tget_class=0, H5Tget_precision=16, H5Tget_size=2)
        else:
            props.update(H5Tget_class=1, H5Tget_precision=32, H5Tget_size=4)
        return props

    def get_bounds(self, chan):
        return (self.first, self.first + self.channels[chan].shape[0])

    def read_vector(self, start, n, chan, sub=None):
        d = self.channels[chan]
        lo = int(start) - self.first
        if lo < 0 or lo + n > d.shape[0]:
            raise IOError("read outside the recording")
        self.reads.append((int(start), int(n)))
        out = d[lo:lo + int(n)]
        if sub is not None and d.ndim == 2:
            out = out[:, sub]
        return out.astype(np.complex64)

    def read_vector_raw(self, start, n, chan, sub=None):
        """Samples as stored: Digital RF keeps complex integers as a structured ('r', 'i') dtype."""
        d = self.channels[chan]
        lo = int(start) - self.first
        if lo < 0 or lo + n > d.shape[0]:
            raise IOError("read outside the recording")
        self.reads.append((int(start), int(n)))
        out = d[lo:lo + int(n)]
        if sub is not None and d.ndim == 2:
            out = out[:, sub]
        if not self.int16:
            return out.astype(np.complex64)
        raw = np.empty(out.shape, dtype=np.dtype([("r", np.int16), ("i", np.int16)]))
        raw["r"] = np.round(out.real).astype(np.int16)
        raw["i"] = np.round(out.imag).astype(np.int16)
        return raw
