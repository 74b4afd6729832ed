"""Epoch loop shared by the algorithms: collect -> update_per_epoch -> eval -> snapshot.

Semantics of reference torchrl/algo/rl_algo.py:19-199 (constructor arguments, timing counters,
best/periodic/final snapshots written as model_<name>_<tag>.pth + _obs_normalizer_<tag>.pkl so
the reference viewers keep loading them).  Not part of the accelerated path: it only drives
the collector (CPU simulator side) and calls update_per_epoch().
"""
import os
import pathlib
import pickle
import time
from collections import deque

import numpy as np
import torch


def _is_box(space):
  return type(space).__name__ == "Box"


class RLAlgo:
  def __init__(self, env=None, replay_buffer=None, collector=None, logger=None, grad_clip=None,
               discount=0.99, num_epochs=3000, batch_size=128, device="cpu", save_interval=100,
               eval_interval=1, save_dir=None):
    self.env = env
    self.continuous = _is_box(self.env.action_space)
    self.replay_buffer = replay_buffer
    self.collector = collector
    self.device = device
    self.discount = discount
    self.num_epochs = num_epochs
    self.epoch_frames = self.collector.epoch_frames
    self.batch_size = batch_size
    self.training_update_num = 0
    self.sample_key = None
    self.grad_clip = grad_clip
    self.logger = logger
    self.episode_rewards = deque(maxlen=30)
    self.training_episode_rewards = deque(maxlen=30)
    self.save_interval = save_interval
    self.save_dir = save_dir
    pathlib.Path(self.save_dir).mkdir(parents=True, exist_ok=True)
    self.best_eval = None
    self.eval_interval = eval_interval
    self.explore_time = 0
    self.train_time = 0
    self.start = time.time()

  # hooks
  def start_epoch(self):
    pass

  def finish_epoch(self):
    return {}

  def pretrain(self):
    pass

  def update_per_epoch(self):
    pass

  def update(self, batch):
    raise NotImplementedError

  def snapshot(self, prefix, epoch):
    normalizer = getattr(self.env, "_obs_normalizer", None)
    if normalizer is not None:
      with open(os.path.join(prefix, "_obs_normalizer_{}.pkl".format(epoch)), "wb") as f:
        pickle.dump(normalizer, f)
    for name, network in self.snapshot_networks:
      # parameters may be views into one flat bucket (PPO engine): clone so that each file holds only
      # its own tensors, with independent storages, like the reference's checkpoints (rl_algo.py:84-95)
      sd = {k: v.detach().clone() for k, v in network.state_dict().items()}
      torch.save(sd, os.path.join(prefix, "model_{}_{}.pth".format(name, epoch)))

  def train(self):
    self.pretrain()
    total_frames = getattr(self, "pretrain_frames", 0)
    self.start_epoch()
    for epoch in range(self.num_epochs):
      self.current_epoch = epoch
      self.start_epoch()

      t0 = time.time()
      training_epoch_info = self.collector.train_one_epoch()
      self.training_episode_rewards.extend(training_epoch_info["train_rewards"])
      self.explore_time += time.time() - t0

      t0 = time.time()
      self.update_per_epoch()
      self.train_time += time.time() - t0

      finish_epoch_info = self.finish_epoch()
      total_frames += self.epoch_frames

      if epoch % self.eval_interval == 0:
        t0 = time.time()
        eval_infos = self.collector.eval_one_epoch()
        eval_time = time.time() - t0
        rewards = eval_infos.pop("eval_rewards")
        self.episode_rewards.extend(rewards)
        mean_eval = np.mean(rewards)
        if self.best_eval is None or mean_eval > self.best_eval:
          self.best_eval = mean_eval
          self.snapshot(self.save_dir, "best")
          print("Best Saved: {:.5f},  EPoch: {}".format(mean_eval, epoch))
        infos = {
          "Running_Average_Rewards": np.mean(self.episode_rewards),
          "Train_Epoch_Reward": training_epoch_info["train_epoch_reward"],
          "Running_Training_Average_Rewards": np.mean(self.training_episode_rewards),
          "Explore_Time": self.explore_time,
          "Train___Time": self.train_time,
          "Eval____Time": eval_time,
        }
        self.explore_time = 0
        self.train_time = 0
        infos.update(eval_infos)
        infos.update(finish_epoch_info)
        self.logger.add_epoch_info(epoch, total_frames, time.time() - self.start, infos)
        self.start = time.time()

      if epoch % self.save_interval == 0:
        self.snapshot(self.save_dir, epoch)

    self.snapshot(self.save_dir, "finish")
    self.collector.terminate()

  @property
  def networks(self):
    return []

  @property
  def snapshot_networks(self):
    return []

  @property
  def target_networks(self):
    return []

  def to(self, device):
    for net in self.networks:
      net.to(device)
