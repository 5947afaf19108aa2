"""tools/eval_cnn.py restates the reference's CNN actor (policy/_network.py, policy/actor_critic.py:240-284) so that
BASELINE configs[4] can run on the GPU box, where the reference tree does not exist.  Where the reference IS present
(the authoring container) the restatement is held to it: same parameter names (a reference state dict loads), same
logits.  Caller context, not product -- but the evaluation numbers quote it."""

from __future__ import annotations

import os
import sys
import types

import pytest
import torch

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load_tool():
    path = os.path.join(os.path.dirname(HERE), "tools", "eval_cnn.py")
    src = open(path).read().replace("import ml2048_b200\n", "")  # the networks need neither CUDA nor the extension
    mod = types.ModuleType("eval_cnn_under_test")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not mounted (GPU box)")
def test_actor_restatement_matches_the_reference_policy():
    sys.path.insert(0, REF_SRC)
    try:
        from ml2048.policy.actor_critic import CNNActorCriticPolicy
    finally:
        sys.path.remove(REF_SRC)
    tool = _load_tool()
    torch.manual_seed(3)
    ref = CNNActorCriticPolicy(share_encoder=True).eval()
    mine = tool.ActorPolicy().eval()
    missing, unexpected = mine.load_state_dict(ref.state_dict(), strict=False)
    assert not missing
    assert all(k.startswith("_critic.") for k in unexpected), unexpected  # the evaluation only needs the actor
    boards = torch.randint(0, 16, (513, 16))
    valid = torch.ones((513, 4), dtype=torch.bool)
    with torch.no_grad():
        want = ref.action_logits(boards, valid)
        onehot = torch.nn.functional.one_hot(boards, 16).float().permute(0, 2, 1).contiguous()  # policy/_network.py:86-95
        got = mine(onehot)
    torch.testing.assert_close(got, want, rtol=0, atol=1e-6)  # fp32, same op order up to the fused input encoding


def test_actor_restatement_shapes_without_the_reference():
    tool = _load_tool()
    torch.manual_seed(0)
    net = tool.ActorPolicy().eval()
    x = torch.zeros((7, 16, 16))
    x[:, 0, :] = 1.0  # empty boards
    with torch.no_grad():
        logits = net(x)
    assert logits.shape == (7, 4) and float(logits.max()) == 0.0  # translated so that max = 0 (policy/_network.py:181-185)
