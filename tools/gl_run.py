import sys; sys.path.insert(0, "visual-context-attentional-gan_b200")
import torch
from vcagan_b200 import audio
spec = torch.rand(64, 321, 300, device="cuda")
w = audio.griffin_lim(spec, None, 3); torch.cuda.synchronize(); print(w.shape)
