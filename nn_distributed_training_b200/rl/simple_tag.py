"""Batched predator-prey environment (MPE ``simple_tag``, continuous actions).

The reference vendors PettingZoo 1.10's pure-Python/numpy MPE stack (RL/pettingzoo/**, ~3.9k
lines) with two local edits to ``simple_tag`` — fixed obstacle positions and an observation
without landmark positions (RL/pettingzoo/mpe/scenarios/simple_tag.py:50-57,134-151) — and steps
ONE environment one agent at a time from Python (RL/dist_rl/dist_ppo.py:200-271).  This module
is a from-scratch tensor implementation of that modified scenario: ``E`` independent worlds are
stepped together with array ops on any torch device, so rollouts are batched matmuls instead
of a per-agent Python loop.

Physics (RL/pettingzoo/mpe/_mpe_utils/core.py): point masses, dt 0.1, damping 0.25, soft
contact force k=1e2 with margin 1e-3 between every colliding pair, per-agent max speed.
Within one cycle every agent observes the same pre-step world state, so the AEC turn order of
the original is equivalent to the simultaneous (parallel) step used here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import torch

OBSTACLE_POSITIONS = [(-1.2, -0.6), (0.1, -1.1), (-0.3, 0.4), (0.9, 0.75),
                      (-0.9, 1.2), (-0.1, 1.3), (-1.2, 0.0), (1.3, 0.0)]


@dataclass
class Space:
    shape: Tuple[int, ...]


class SimpleTagEnv:
    """``num_adversaries`` predators (agents 0..A-1) chase ``num_good`` prey (last agents)."""

    DT, DAMPING, CONTACT_FORCE, CONTACT_MARGIN = 0.1, 0.25, 1e2, 1e-3

    def __init__(self, num_envs=1, num_good=1, num_adversaries=3, num_obstacles=2, max_cycles=25,
                 device="cpu", dtype=torch.float32, seed=None):
        if num_obstacles > len(OBSTACLE_POSITIONS):
            raise ValueError("at most 8 obstacles (fixed positions)")
        self.E, self.n_adv, self.n_good, self.n_obs = num_envs, num_adversaries, num_good, num_obstacles
        self.A = num_adversaries + num_good
        self.max_cycles = max_cycles
        self.device, self.dtype = torch.device(device), dtype
        self.gen = torch.Generator(device="cpu")
        if seed is not None:
            self.gen.manual_seed(seed)
        kw = dict(device=self.device, dtype=dtype)
        adv = torch.arange(self.A) < num_adversaries
        self.is_adv = adv.to(self.device)
        self.size = torch.where(adv, 0.075, 0.05).to(**kw)
        self.accel = torch.where(adv, 3.0, 4.0).to(**kw)
        self.max_speed = torch.where(adv, 1.0, 1.3).to(**kw)
        self.obst = torch.tensor(OBSTACLE_POSITIONS[:num_obstacles], **kw).reshape(num_obstacles, 2)
        self.obst_size = 0.2
        self.agents = [f"adversary_{i}" for i in range(num_adversaries)] + [f"agent_{i}" for i in range(num_good)]
        obs_adv = 4 + 2 * (self.A - 1) + 2 * num_good
        obs_good = 4 + 2 * (self.A - 1) + 2 * (num_good - 1)
        self.observation_spaces: Dict[str, Space] = {n: Space((obs_adv if n.startswith("adv") else obs_good,)) for n in self.agents}
        self.action_spaces: Dict[str, Space] = {n: Space((5,)) for n in self.agents}
        self.num_agents = self.A
        self.reset()

    # ------------------------------------------------------------------
    def reset(self):
        kw = dict(device=self.device, dtype=self.dtype)
        self.pos = (torch.rand(self.E, self.A, 2, generator=self.gen) * 2 - 1).to(**kw)
        self.vel = torch.zeros(self.E, self.A, 2, **kw)
        self.cycle = 0
        return self.observe()

    def observe(self):
        """(obs_adversaries [E, n_adv, 4+2(A-1)+2 n_good], obs_good [E, n_good, ...]) laid out as in the
        reference scenario: own vel, own pos, relative positions of the other agents (index order),
        velocities of the good agents among the others."""
        rel = self.pos.unsqueeze(1) - self.pos.unsqueeze(2)          # rel[e, i, j] = pos_j - pos_i
        out = []
        for i in range(self.A):
            others = [j for j in range(self.A) if j != i]
            parts = [self.vel[:, i], self.pos[:, i], rel[:, i, others].reshape(self.E, -1)]
            good_others = [j for j in others if j >= self.n_adv]
            if good_others:
                parts.append(self.vel[:, good_others].reshape(self.E, -1))
            out.append(torch.cat(parts, dim=1))
        obs_adv = torch.stack(out[: self.n_adv], dim=1)
        obs_good = torch.stack(out[self.n_adv:], dim=1)
        return obs_adv, obs_good

    # ------------------------------------------------------------------
    def _contact_forces(self):
        k = self.CONTACT_MARGIN
        # agent-agent
        d = self.pos.unsqueeze(2) - self.pos.unsqueeze(1)             # d[e,a,b] = pos_a - pos_b
        dist = d.norm(dim=-1).clamp_min(1e-12)
        dmin = self.size.view(1, -1, 1) + self.size.view(1, 1, -1)
        pen = torch.nn.functional.softplus(-(dist - dmin) / k) * k
        f = self.CONTACT_FORCE * d / dist.unsqueeze(-1) * pen.unsqueeze(-1)
        eye = torch.eye(self.A, device=self.device, dtype=torch.bool).view(1, self.A, self.A, 1)
        force = f.masked_fill(eye, 0.0).sum(dim=2)
        # agent-obstacle (obstacles are immovable)
        if self.n_obs:
            d = self.pos.unsqueeze(2) - self.obst.view(1, 1, -1, 2)
            dist = d.norm(dim=-1).clamp_min(1e-12)
            dmin = self.size.view(1, -1, 1) + self.obst_size
            pen = torch.nn.functional.softplus(-(dist - dmin) / k) * k
            force = force + (self.CONTACT_FORCE * d / dist.unsqueeze(-1) * pen.unsqueeze(-1)).sum(dim=2)
        return force

    def step(self, actions: torch.Tensor):
        """``actions [E, A, 5]`` = (noop, +x, -x, +y, -y) intensities.  Returns (rew_adv [E,n_adv],
        rew_good [E,n_good], done)."""
        a = actions.to(device=self.device, dtype=self.dtype)
        u = torch.stack([a[..., 1] - a[..., 2], a[..., 3] - a[..., 4]], dim=-1) * self.accel.view(1, -1, 1)
        force = u + self._contact_forces()
        vel = self.vel * (1 - self.DAMPING) + force * self.DT         # unit mass
        speed = vel.norm(dim=-1, keepdim=True)
        ms = self.max_speed.view(1, -1, 1)
        vel = torch.where(speed > ms, vel / speed.clamp_min(1e-12) * ms, vel)
        self.vel = vel
        self.pos = self.pos + vel * self.DT
        self.cycle += 1
        return (*self._rewards(), self.cycle >= self.max_cycles)

    # ------------------------------------------------------------------
    def render(self, world: int = 0, size: int = 400, pos=None):
        """RGB frame of one world (the reference renders through pyglet, _mpe_utils/rendering.py): predators red,
        prey green, obstacles grey, view [-1.3, 1.3]^2.  ``pos`` overrides the agent positions (recorded rollouts)."""
        from PIL import Image, ImageDraw
        im = Image.new("RGB", (size, size), (255, 255, 255))
        d = ImageDraw.Draw(im)
        cam = 1.3

        def circle(xy, r, fill):
            cx, cy = (xy[0] + cam) / (2 * cam) * size, (cam - xy[1]) / (2 * cam) * size
            rr = r / (2 * cam) * size
            d.ellipse([cx - rr, cy - rr, cx + rr, cy + rr], fill=fill, outline=(0, 0, 0))
        for o in self.obst.tolist():
            circle(o, self.obst_size, (64, 64, 64))
        p = self.pos[world].tolist() if pos is None else [list(map(float, q)) for q in pos]
        for i, q in enumerate(p):
            circle(q, float(self.size[i]), (217, 89, 89) if i < self.n_adv else (89, 217, 89))
        return im

    def _rewards(self):
        adv, good = self.pos[:, : self.n_adv], self.pos[:, self.n_adv:]
        dist = (adv.unsqueeze(2) - good.unsqueeze(1)).norm(dim=-1)    # [E, n_adv, n_good]
        coll = dist < (0.075 + 0.05)
        # shaped, shared adversary reward: -0.1 sum_adv min_good dist + 10 per colliding pair
        r_adv = -0.1 * dist.min(dim=2).values.sum(dim=1) + 10.0 * coll.sum(dim=(1, 2)).to(self.dtype)
        r_adv = r_adv.unsqueeze(1).expand(-1, self.n_adv)
        x = good.abs()
        bound = torch.where(x < 0.9, torch.zeros_like(x),
                            torch.where(x < 1.0, (x - 0.9) * 10, torch.exp(2 * x - 2).clamp_max(10.0)))
        r_good = -10.0 * coll.sum(dim=1).to(self.dtype) - bound.sum(dim=-1)
        return r_adv, r_good


def heuristic_prey_action(obs_good: torch.Tensor, n_adv: int) -> torch.Tensor:
    """Scripted evader: move away from the closest predator, never past |x| = 1.2
    (reference: RL/dist_rl/dist_ppo.py:79-126).  ``obs_good [E, 4 + 2 n_adv]`` -> actions ``[E, 5]``."""
    rel = obs_good[:, 4: 4 + 2 * n_adv].reshape(-1, n_adv, 2)
    near = rel.norm(dim=-1).argmin(dim=1)
    d = rel[torch.arange(rel.shape[0], device=rel.device), near]
    force = -d / d.abs().max(dim=1, keepdim=True).values.clamp_min(1e-12)
    act = torch.zeros(obs_good.shape[0], 5, device=obs_good.device, dtype=obs_good.dtype)
    act[:, 1] = force[:, 0].clamp_min(0)
    act[:, 2] = (-force[:, 0]).clamp_min(0)
    act[:, 3] = force[:, 1].clamp_min(0)
    act[:, 4] = (-force[:, 1]).clamp_min(0)
    px, py = obs_good[:, 2], obs_good[:, 3]
    act[:, 2] = torch.where(px <= -1.2, torch.zeros_like(px), act[:, 2])
    act[:, 1] = torch.where(px >= 1.2, torch.zeros_like(px), act[:, 1])
    act[:, 4] = torch.where(py <= -1.2, torch.zeros_like(py), act[:, 4])
    act[:, 3] = torch.where(py >= 1.2, torch.zeros_like(py), act[:, 3])
    return act
