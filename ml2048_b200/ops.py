"""Stand-alone device ops of the rollout path (thin wrappers over the C ABI, torch CUDA tensors in and out).

    sample_masked_categorical   _sample_action, policy/actor_critic.py:56-76
    gae_advantages              the recurrence of compute_gae, gae.py:50, :65-68
    encode_onehot               CNNEncoder.forward's input encoding, policy/_network.py:86-95
    valid_actions               _compute_valid_actions, game_numba.py:259-289
"""

from __future__ import annotations

from typing import Optional

import torch

from . import _lib

_ONEHOT = {torch.float32: _lib.ONEHOT_F32, torch.bfloat16: _lib.ONEHOT_BF16, torch.uint8: _lib.ONEHOT_U8}


def _stream(t: torch.Tensor) -> int:
    from .vecgame import _raw_stream

    return _raw_stream(t.device.index)


def _cuda(t: torch.Tensor, name: str) -> None:
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous CUDA tensor")


def sample_masked_categorical(logits: torch.Tensor, valid: torch.Tensor, *, seed: int, counter: int, slot_base: int = 0,
                              dtype: torch.dtype = torch.int64) -> tuple[torch.Tensor, torch.Tensor]:
    """Sample one action per game from softmax(logits restricted to valid) and return (actions, log_probs).

    Same distribution and log-probabilities as the reference's ``_sample_action`` (invalid actions masked to
    finfo.min, ``Categorical(logits=...)``); the uniform comes from Philox2x32-10 keyed by
    (seed, slot_base + game, counter) -- the policy word the step kernel itself draws -- instead of ``torch.multinomial``'s generator."""
    _cuda(logits, "logits")
    _cuda(valid, "valid")
    if logits.dtype != torch.float32 or logits.ndim != 2 or logits.shape[1] != 4:
        raise ValueError(f"logits must be float32 (M,4), got {logits.dtype}{tuple(logits.shape)}")
    if valid.shape != logits.shape or valid.dtype not in (torch.bool, torch.uint8):
        raise ValueError(f"valid must be bool/uint8 {tuple(logits.shape)}")
    if dtype not in (torch.int64, torch.uint8):
        raise ValueError("dtype must be torch.int64 or torch.uint8")
    m = logits.shape[0]
    actions = torch.empty((m,), dtype=dtype, device=logits.device)
    log_prob = torch.empty((m,), dtype=torch.float32, device=logits.device)
    lib = _lib.load()
    with torch.cuda.device(logits.device):
        rc = lib.ml2048_sample_masked_categorical(
            logits.data_ptr(), valid.data_ptr(), actions.data_ptr() if dtype == torch.uint8 else None,
            actions.data_ptr() if dtype == torch.int64 else None, log_prob.data_ptr(), m, int(slot_base),
            int(seed) & 0xFFFFFFFFFFFFFFFF, int(counter) & 0xFFFFFFFFFFFFFFFF, _stream(logits))
    _lib.check(rc, "ml2048_sample_masked_categorical")
    return actions, log_prob


def gae_advantages(v0: torch.Tensor, v1: torch.Tensor, reward: torch.Tensor, terminated: torch.Tensor, *, gamma: float,
                   lambda_: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Generalised advantage estimation over (use, step, game) tensors, as ``compute_gae`` (gae.py:50, :65-68):
    ``delta = gamma*v1*~terminated + reward - v0`` then the reverse scan with ``coef = gamma*lambda_``.
    One kernel instead of 3 torch ops per step; same fp32 rounding as the reference's ops."""
    for name, t in (("v0", v0), ("v1", v1), ("reward", reward), ("terminated", terminated)):
        _cuda(t, name)
    if v0.ndim != 3 or v0.dtype != torch.float32:
        raise ValueError(f"v0 must be float32 (use, step, game), got {v0.dtype}{tuple(v0.shape)}")
    if v1.shape != v0.shape or reward.shape != v0.shape or terminated.shape != v0.shape:
        raise ValueError("v0, v1, reward, terminated must have the same shape")
    if v1.dtype != torch.float32 or reward.dtype != torch.float32 or terminated.dtype not in (torch.bool, torch.uint8):
        raise ValueError("v1/reward must be float32 and terminated bool/uint8")
    if out is None:
        out = torch.empty_like(v0)
    _cuda(out, "out")
    if out.shape != v0.shape or out.dtype != torch.float32:
        raise ValueError("out must be float32 with the shape of v0")
    u, s, g = v0.shape
    lib = _lib.load()
    with torch.cuda.device(v0.device):
        rc = lib.ml2048_gae(v0.data_ptr(), v1.data_ptr(), reward.data_ptr(), terminated.data_ptr(), out.data_ptr(), u, s, g,
                            float(gamma), float(gamma * lambda_), _stream(v0))
    _lib.check(rc, "ml2048_gae")
    return out


def encode_onehot(board: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """(M,16) uint8 boards -> (M,16,16) class-major one-hot == F.one_hot(x,16).to(dtype).permute(0,2,1)."""
    _cuda(board, "board")
    if board.dtype not in (torch.uint8, torch.int8) or board.ndim != 2 or board.shape[1] != 16:
        raise ValueError("board must be uint8/int8 (M,16)")
    if dtype not in _ONEHOT:
        raise ValueError(f"dtype {dtype} not supported")
    out = torch.empty((board.shape[0], 16, 16), dtype=dtype, device=board.device)
    lib = _lib.load()
    with torch.cuda.device(board.device):
        rc = lib.ml2048_encode_onehot(board.data_ptr(), out.data_ptr(), _ONEHOT[dtype], board.shape[0], _stream(board))
    _lib.check(rc, "ml2048_encode_onehot")
    return out


def valid_actions(board: torch.Tensor) -> torch.Tensor:
    """(M,16) uint8 boards -> (M,4) uint8 masks (left, right, up, down)."""
    _cuda(board, "board")
    if board.dtype not in (torch.uint8, torch.int8) or board.ndim != 2 or board.shape[1] != 16:
        raise ValueError("board must be uint8/int8 (M,16)")
    out = torch.empty((board.shape[0], 4), dtype=torch.uint8, device=board.device)
    lib = _lib.load()
    with torch.cuda.device(board.device):
        rc = lib.ml2048_valid_actions(board.data_ptr(), out.data_ptr(), board.shape[0], _stream(board))
    _lib.check(rc, "ml2048_valid_actions")
    return out
