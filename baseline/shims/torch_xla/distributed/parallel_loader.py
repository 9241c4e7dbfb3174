"""pl.MpDeviceLoader: move each batch to the device."""
import torch


class MpDeviceLoader:
    def __init__(self, loader, device):
        self.loader, self.device = loader, device

    def __iter__(self):
        for batch in self.loader:
            yield tuple(torch.as_tensor(t).to(self.device, non_blocking=True) for t in batch)

    def __len__(self):
        return len(self.loader)
