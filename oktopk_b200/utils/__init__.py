"""Utilities: logging, metrics/timers/NVTX, CTC decoding + WER, analytic perf models, settings, elastic helpers."""
