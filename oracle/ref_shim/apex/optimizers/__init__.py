import torch
FusedAdam = torch.optim.AdamW
FusedSGD = torch.optim.SGD
