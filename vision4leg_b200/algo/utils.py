"""Schedule / target-copy helpers (reference torchrl/algo/utils.py:16-32)."""


def soft_update_from_to(source, target, tau):
  for t, s in zip(target.parameters(), source.parameters()):
    t.data.mul_(1.0 - tau).add_(s.data, alpha=tau)


def copy_model_params_from_to(source, target):
  for t, s in zip(target.parameters(), source.parameters()):
    t.data.copy_(s.data)


def linear_lr(epoch, total_num_epochs, initial_lr):
  """lr = lr0 - lr0 * epoch / num_epochs (utils.py:28-32)"""
  return initial_lr - (initial_lr * (epoch / float(total_num_epochs)))


def update_linear_schedule(optimizer, epoch, total_num_epochs, initial_lr):
  lr = linear_lr(epoch, total_num_epochs, initial_lr)
  for group in optimizer.param_groups:
    group["lr"] = lr
