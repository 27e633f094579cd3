"""Small algorithm helpers (API of /root/reference/torchrl/algo/utils.py:5-32)."""
import torch


def huber(x, k=1.0):
    return torch.where(x.abs() < k, 0.5 * x.pow(2), k * (x.abs() - 0.5 * k))


def quantile_regression_loss(coefficient, source, target):
    """Plain-torch statement of the QR loss (utils.py:5-9); the hot path uses ops.qr_huber_loss."""
    diff = target.unsqueeze(-1) - source.unsqueeze(1)
    loss = huber(diff) * (coefficient - (diff.detach() < 0).float()).abs()
    return loss.mean()


def soft_update_from_to(source, target, tau):
    """theta' <- (1-tau) theta' + tau theta, parameter by parameter (cold path; agents with flat
    buffers use ops.polyak_update on the whole buffer in one launch)."""
    with torch.no_grad():
        for tp, p in zip(target.parameters(), source.parameters()):
            tp.data.mul_(1.0 - tau).add_(p.data, alpha=tau)


def copy_model_params_from_to(source, target):
    with torch.no_grad():
        for tp, p in zip(target.parameters(), source.parameters()):
            tp.data.copy_(p.data)


def update_linear_schedule(optimizer, epoch, total_num_epochs, initial_lr):
    """Linear LR decay (utils.py:28-32)."""
    lr = initial_lr - (initial_lr * (epoch / float(total_num_epochs)))
    for param_group in optimizer.param_groups:
        param_group['lr'] = lr
