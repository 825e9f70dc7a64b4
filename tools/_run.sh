python tools/latency_probe.py
