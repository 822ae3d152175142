"""Geodesic pose loss -- mirror of /root/reference/src/geom/losses.py:3-21 on top of the SE3 kernels."""
import torch


_IDX = {}


def _pair_indices(device):
    """The reference builds torch.as_tensor([0, 1]) / [1, 0] on every call (losses.py:5-6): a pageable host -> device copy
    per step, which also cannot be recorded in a CUDA graph.  Same tensors, created once per device."""
    key = str(device)
    if key not in _IDX:
        _IDX[key] = (torch.tensor([0, 1], device=device), torch.tensor([1, 0], device=device))
    return _IDX[key]


def geodesic_loss(Ps, Gs, train_val="train", sync_metrics=True):
    """Ps: ground-truth SE3 [B,2]; Gs: list holding the predicted SE3 [B,2] (model output).
    d = log((G_j G_i^-1) (P_j P_i^-1)^-1) for (i,j) in ((0,1),(1,0)); mean |tau|, mean |phi|.
    The reference turns both losses into Python floats (`.item()`, losses.py:17-18): a device->host sync per step.
    sync_metrics=False keeps them as 0-d device tensors for callers that log asynchronously."""
    ii, jj = _pair_indices(Ps.data.device)
    dP = Ps[:, jj] * Ps[:, ii].inv()
    dG = Gs[0][:, jj] * Gs[0][:, ii].inv()
    d = (dG * dP.inv()).log()
    tau, phi = d.split([3, 3], dim=-1)
    loss_tr = tau.norm(dim=-1).mean()
    loss_rot = phi.norm(dim=-1).mean()
    if sync_metrics:
        metrics = {
            train_val + "_geo_loss_tr": loss_tr.detach().item(),
            train_val + "_geo_loss_rot": loss_rot.detach().item(),
        }
    else:
        metrics = {train_val + "_geo_loss_tr": loss_tr.detach(), train_val + "_geo_loss_rot": loss_rot.detach()}
    return loss_tr, loss_rot, metrics
