"""Synthetic Atari-shaped pixel env -- CPU definition, batched NumPy, pure integer arithmetic
(TEST INFRASTRUCTURE).  BASELINE.json config 4 asks for a "synthetic pixel env (4x84x84, 6 actions)"; the
reference only wraps real ALE games (/root/reference/torchrl/env/atari_wrapper.py), so the game is defined
by this build: a ball bouncing in an 84x84 field and a paddle the agent moves; observation = the last 4
rendered uint8 frames (the shape WarpFrame + FrameStack produce, atari_wrapper.py:112-168).  PARITY UNPINNED
by the reference; pinned GPU-vs-this-file BIT-EXACTLY (everything is integer).

  latent   bx, by (ball top-left), vx, vy, px (paddle left edge)
  actions  0 NOOP, 1 FIRE(noop), 2 RIGHT +3, 3 LEFT -3, 4 RIGHTFIRE +6, 5 LEFTFIRE -6      (px clipped to [0,72])
  step     ball moves by (vx,vy); reflects off the side/top walls; when its bottom reaches the paddle row
           (by >= 74): hit  (px-3 <= bx <= px+11) -> reward +1, reflect up;  miss -> reward -1, done
           time limit: 1000 steps -> done with the time_limit flag
  render   background ((7x + 13y) & 31) + 16, ball 4x4 of 255, paddle 12x2 of 200 at rows 78..79
  reset    latent from the murmur3-finaliser hash of (seed, episode, j) (same hash as oracle/synth_env.py)
"""
import numpy as np

from oracle.synth_env import mix32

H = W = 84
PADDLE_Y, PADDLE_W, BALL = 78, 12, 4
MAX_STEPS = 1000
_M32 = 0xFFFFFFFF


def _hash(seed, episode, j):
    key = (seed * 0x9E3779B1 + episode * 0x85EBCA77 + j * 0xC2B2AE3D + 0x27D4EB2F) & _M32
    return mix32(key)


def reset_latent(seeds, episodes):
    """(n,) uint64 seeds / episodes -> dict of int64 arrays."""
    s = np.asarray(seeds, dtype=np.uint64)
    e = np.asarray(episodes, dtype=np.uint64)
    h = [np.asarray(_hash(s, e, np.uint64(j)), dtype=np.int64) for j in range(5)]
    bx = 4 + h[0] % 72
    by = 4 + h[1] % 40
    vx = np.where(h[2] & 1, 1, -1) * (1 + ((h[2] >> 1) & 1))
    vy = 1 + (h[3] & 1)
    px = h[4] % 73
    return {"bx": bx, "by": by, "vx": vx, "vy": vy, "px": px}


_DX = np.array([0, 0, 3, -3, 6, -6], dtype=np.int64)


def step_latent(lat, actions):
    """One transition.  Returns (new latent, reward int64 (n,), done_dyn bool (n,))."""
    a = np.asarray(actions).astype(np.int64)
    px = np.clip(lat["px"] + _DX[a], 0, W - PADDLE_W)
    bx = lat["bx"] + lat["vx"]
    by = lat["by"] + lat["vy"]
    vx, vy = lat["vx"].copy(), lat["vy"].copy()
    lo = bx < 0
    bx = np.where(lo, -bx, bx); vx = np.where(lo, -vx, vx)
    hi = bx > W - BALL
    bx = np.where(hi, 2 * (W - BALL) - bx, bx); vx = np.where(hi, -vx, vx)
    top = by < 0
    by = np.where(top, -by, by); vy = np.where(top, -vy, vy)
    at_paddle = by >= PADDLE_Y - BALL
    hit = at_paddle & (bx >= px - 3) & (bx <= px + PADDLE_W - 1)
    miss = at_paddle & ~hit
    by = np.where(hit, 2 * (PADDLE_Y - BALL) - by, by)
    vy = np.where(hit, -vy, vy)
    reward = hit.astype(np.int64) - miss.astype(np.int64)
    return {"bx": bx, "by": by, "vx": vx, "vy": vy, "px": px}, reward, miss


def render(lat):
    """(n, 84, 84) uint8 frames."""
    n = len(lat["bx"])
    y, x = np.mgrid[0:H, 0:W]
    bg = (((7 * x + 13 * y) & 31) + 16).astype(np.uint8)
    fr = np.broadcast_to(bg, (n, H, W)).copy()
    X, Y = x[None], y[None]
    bx, by, px = (lat[k].reshape(-1, 1, 1) for k in ("bx", "by", "px"))
    paddle = (Y >= PADDLE_Y) & (Y < PADDLE_Y + 2) & (X >= px) & (X < px + PADDLE_W)
    fr[paddle] = 200
    ball = (X >= bx) & (X < bx + BALL) & (Y >= by) & (Y < by + BALL)
    fr[ball] = 255
    return fr
