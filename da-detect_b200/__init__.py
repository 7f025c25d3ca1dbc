"""dadetect_b200 — Blackwell (sm_100a) native training hot path for the Domain-Adaptive
Faster R-CNN of jinlong17/DA-Detect.  See DESIGN.md."""
__version__ = "0.1.0"
