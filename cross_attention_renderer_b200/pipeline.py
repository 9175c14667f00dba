"""Host-to-device pipelining for render loops whose scenes live in (pinned) host memory.

The reference's drivers upload one scene, render it chunk by chunk and download the image before touching the next
scene (render_realestate10k_traj.py:84-150, eval_realestate10k.py:131-199).  ``render_host_batches`` keeps the same
per-batch semantics - every batch's inputs cross PCIe, every batch's rgb / valid_mask / depth_ray come back to the
host - but uploads batch k+1 on a copy stream while batch k renders, and only blocks the host on batch k's results."""
import torch


def _upload(obj, dev):
    if isinstance(obj, dict):
        return {k: _upload(v, dev) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_upload(v, dev) for v in obj)
    if torch.is_tensor(obj):
        return obj.to(dev, non_blocking=True)
    return obj


def _record(obj, stream):
    if isinstance(obj, dict):
        for v in obj.values():
            _record(v, stream)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record(v, stream)
    elif torch.is_tensor(obj) and obj.is_cuda:
        obj.record_stream(stream)


def render_host_batches(model, batches, device, keys=("rgb", "valid_mask", "depth_ray")):
    """``batches``: iterable of ``(input_dict, z_list)`` on the host (pin them for asynchronous copies).
    Yields ``{key: host tensor}`` per batch, in order."""
    dev = torch.device(device)
    main = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(dev)

    def start(batch):
        inp_h, z_h = batch
        with torch.cuda.stream(copy):
            inp_d, z_d = _upload(inp_h, dev), _upload(z_h, dev)
            ev = torch.cuda.Event()
            ev.record(copy)
        return inp_d, z_d, ev

    it = iter(batches)
    try:
        nxt = start(next(it))
    except StopIteration:
        return
    while nxt is not None:
        inp_d, z_d, ev = nxt
        try:
            nxt = start(next(it))                      # in flight while this batch renders
        except StopIteration:
            nxt = None
        main.wait_event(ev)
        _record(inp_d, main)
        _record(z_d, main)
        with torch.no_grad():
            out = model(inp_d, z=z_d)
        yield {k: out[k].cpu() for k in keys}          # D2H; blocks the host on this batch only
