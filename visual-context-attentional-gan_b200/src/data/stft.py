"""Drop-in for the reference's src/data/stft.py (`from src.data.stft import STFT`): CUDA STFT/ISTFT (csrc/stft.cu)."""
from vcagan_b200.audio import STFT  # noqa: F401
