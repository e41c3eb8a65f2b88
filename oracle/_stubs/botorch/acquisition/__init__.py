import torch


class AcquisitionFunction(torch.nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model
