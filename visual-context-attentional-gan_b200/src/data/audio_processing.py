"""Drop-in for the reference's src/data/audio_processing.py::griffin_lim (called from vid_aud_grid.py:203,217)."""
from vcagan_b200.audio import griffin_lim  # noqa: F401
