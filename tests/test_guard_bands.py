"""Out-of-bounds check without compute-sanitizer (closed on this pool): every buffer handed to the C ABI is carved
out of one arena with canary-filled gaps on both sides; after running the kernels on awkward sizes the canaries must
be intact and the results must equal a run on ordinary tensors."""

from __future__ import annotations

import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

GAP = 4096
CANARY = 0xA5


class Arena:
    def __init__(self, nbytes: int):
        self.buf = torch.full((nbytes,), CANARY, dtype=torch.uint8, device="cuda")
        self.off = GAP
        self.spans = []

    def take(self, nbytes: int, dtype=torch.uint8, fill=0) -> torch.Tensor:
        start = (self.off + 511) // 512 * 512
        t = self.buf[start:start + nbytes]
        t.fill_(fill)
        self.spans.append((start, start + nbytes))
        self.off = start + nbytes + GAP
        assert self.off < self.buf.numel()
        return t.view(dtype)

    def check(self) -> None:
        mask = torch.ones_like(self.buf, dtype=torch.bool)
        for a, b in self.spans:
            mask[a:b] = False
        bad = (self.buf[mask] != CANARY).nonzero()
        assert bad.numel() == 0, f"{bad.numel()} canary bytes overwritten, first at masked index {int(bad[0])}"


@pytest.mark.parametrize("m", [1, 31, 257, 4095, 4097, 65537])
@pytest.mark.parametrize("onehot", [0, 1, 2, 3])
def test_kernels_stay_inside_their_buffers(m, onehot):
    import ml2048_b200
    from ml2048_b200 import _lib

    lib = _lib.load()
    esz = {0: 0, 1: 4, 2: 2, 3: 1}[onehot]
    pad = (m + 15) // 16 * 16
    ar = Arena(m * (16 * 3 + 4 * 2 + 4 * 4 + 2 + 8 + 4 + 256 * esz + 50 + 16 + 4) + pad + 40 * GAP + (1 << 20))
    board = [ar.take(16 * m), ar.take(16 * m)]
    valid = [ar.take(4 * m), ar.take(4 * m)]
    ids, reward = ar.take(4 * m, torch.int32), ar.take(4 * m, torch.float32)
    step_score = ar.take(8 * m, torch.int32)  # {step, score} records (ml2048_b200.h)
    step_ptr, score_ptr = step_score.data_ptr(), step_score.data_ptr() + 4
    score = step_score.view(m, 2)[:, 1].view(torch.float32)
    term, invalid = ar.take(pad, fill=0), ar.take(m)
    term[:m] = 1
    merged = ar.take(16 * m)
    oh = ar.take(256 * esz * m) if onehot else None
    actions_out, logp = ar.take(m), ar.take(4 * m, torch.float32)
    logits = ar.take(16 * m, torch.float32)
    logits.normal_()
    tr = {"tr_state": ar.take(16 * m), "tr_valid_actions": ar.take(4 * m), "tr_action": ar.take(m), "tr_reward": ar.take(4 * m),
          "tr_next_state": ar.take(16 * m), "tr_next_valid_actions": ar.take(4 * m), "tr_step": ar.take(4 * m), "tr_terminated": ar.take(m)}
    ep = {"episode_steps": ar.take(4 * 64), "episode_score": ar.take(4 * 64), "episode_max_tile": ar.take(64)}
    game_count, reset_count = ar.take(8, torch.int64), ar.take(8, torch.int64)
    indices = ar.take(8 * m, torch.int64)
    scratch = ar.take(4 * int(lib.ml2048_prepare_scratch_ints(m)), torch.int32)
    stats = ar.take(8 * _lib.STATS_WORDS * _lib.STATS_REPLICAS, torch.int64)
    tables = ar.take(2 * 1024 * 16)
    gen = torch.Generator().manual_seed(0)
    perm = torch.stack([torch.randperm(16, generator=gen) for _ in range(1024)]).to(torch.uint8)
    keys = torch.zeros_like(perm)
    assert lib.ml2048_pack_randperm_keys(perm.data_ptr(), keys.data_ptr(), 1024) == 0
    tables.copy_(torch.cat([perm.flatten(), keys.flatten()]).cuda())
    stream = torch.cuda.current_stream().cuda_stream

    # the same run on a regular environment, for the results
    ref = ml2048_b200.VecGame(m, output="torch", sync_free=True, rng_mode="philox")
    ref.reset(9)

    p = _lib.PrepareArgs(struct_size=C.sizeof(_lib.PrepareArgs), rng_mode=_lib.RNG_PHILOX, onehot_dtype=onehot, num_games=m, slot_base=0,
                         id=ids.data_ptr(), step=step_ptr, score=score_ptr, reward=reward.data_ptr(),
                         terminated=term.data_ptr(), invalid=invalid.data_ptr(), merged=merged.data_ptr(),
                         onehot=oh.data_ptr() if onehot else None, randperm=tables.data_ptr(), two_mask=0xFFFF,
                         two_threshold=ref._two_threshold, philox_seed=ref._philox_seed, game_count=game_count.data_ptr(),
                         reset_count=reset_count.data_ptr(), reset_indices=indices.data_ptr(), scratch=scratch.data_ptr())
    a = _lib.StepArgs(struct_size=C.sizeof(_lib.StepArgs), reward_kind=0, rng_mode=_lib.RNG_PHILOX, onehot_dtype=onehot, num_games=m,
                      slot_base=0, step=step_ptr, score=score_ptr, reward=reward.data_ptr(), terminated=term.data_ptr(),
                      invalid=invalid.data_ptr(), merged=merged.data_ptr(), onehot_out=oh.data_ptr() if onehot else None,
                      randperm_keys=tables.data_ptr() + 16384, two_mask=0xFFFF, two_threshold=ref._two_threshold,
                      philox_seed=ref._philox_seed, stats=stats.data_ptr(), actions_out=actions_out.data_ptr(),
                      id=ids.data_ptr(), episode_capacity=64, **{k: v.data_ptr() for k, v in ep.items()})
    cur = 0
    for t in range(40):
        counter = ref._philox_counter
        ref.prepare()
        p.two_mask = a.two_mask = ref._two_mask  # Philox mode: the table epoch's 2-vs-4 mask, drawn on the host per prepare()
        p.board, p.valid, p.philox_counter = board[cur].data_ptr(), valid[cur].data_ptr(), counter
        assert lib.ml2048_prepare(C.byref(p), stream) == 0
        a.board_in, a.board_out = board[cur].data_ptr(), board[1 - cur].data_ptr()
        a.valid_in, a.valid_out = valid[cur].data_ptr(), valid[1 - cur].data_ptr()
        a.philox_counter = counter + 1
        if t % 2:
            a.action_mode, a.logits, a.log_prob_out = _lib.ACTIONS_FROM_LOGITS, logits.data_ptr(), logp.data_ptr()
            for k, v in tr.items():
                setattr(a, k, v.data_ptr())
            ref.step_from_logits(logits.view(m, 4))
        else:
            a.action_mode, a.logits, a.log_prob_out = _lib.ACTIONS_RANDOM_VALID, None, None
            for k in tr:
                setattr(a, k, None)
            ref.step_random()
        assert lib.ml2048_step(C.byref(a), stream) == 0
        cur = 1 - cur
    torch.cuda.synchronize()
    ar.check()
    assert torch.equal(board[cur].view(m, 16), ref.observations()[0])
    assert torch.equal(score, ref._score) and torch.equal(ids, ref._id)
    assert int(game_count[0]) == ref._game_count
