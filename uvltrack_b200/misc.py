"""Small carriers shared by the model and the tracker."""


class NestedTensor:
    """(tensors, mask) pair used for the text ids + attention mask (lib/utils/misc.py:23-46)."""

    def __init__(self, tensors, mask):
        self.tensors = tensors
        self.mask = mask

    def to(self, device):
        return NestedTensor(self.tensors.to(device), None if self.mask is None else self.mask.to(device))

    def decompose(self):
        return self.tensors, self.mask

    def __repr__(self):
        return f"NestedTensor(tensors={tuple(self.tensors.shape)}, mask={None if self.mask is None else tuple(self.mask.shape)})"
