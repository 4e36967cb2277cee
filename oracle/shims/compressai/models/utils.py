"""ORACLE: update_registered_buffers as called at base_model.py:82-94."""
import torch


def _find(module, query):
    return next((b for n, b in module.named_buffers() if n == query), None)


def update_registered_buffers(module, module_name, buffer_names, state_dict,
                              policy="resize_if_empty", dtype=torch.int):
    if not module:
        return
    valid = [n for n, _ in module.named_buffers()]
    for name in buffer_names:
        if name not in valid:
            raise ValueError(f'Invalid buffer name "{name}"')
    for name in buffer_names:
        new_size = state_dict[f"{module_name}.{name}"].size()
        buf = _find(module, name)
        if policy in ("resize_if_empty", "resize"):
            if buf is None:
                raise RuntimeError(f'buffer "{name}" was not registered')
            if policy == "resize" or buf.numel() == 0:
                buf.resize_(new_size)
        elif policy == "register":
            if buf is not None:
                raise RuntimeError(f'buffer "{name}" was already registered')
            module.register_buffer(name, torch.empty(new_size, dtype=dtype).fill_(0))
        else:
            raise ValueError(f'Invalid policy "{policy}"')
