"""Host-side helpers mirroring the reference's `utils/` modules that sit on the hot path
(camera.py:219-323, transforms.py, activations.py, schedulers.py, misc.py)."""
