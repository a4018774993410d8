"""gymtorch: tensors are already torch tensors in the stub."""


def wrap_tensor(t):
    return t


def unwrap_tensor(t):
    return t
