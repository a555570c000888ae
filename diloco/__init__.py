"""Compatibility alias: ``python -m diloco.train @configs/...`` is the entrypoint BASELINE.json names; the engine lives in ``prime_b200``."""
