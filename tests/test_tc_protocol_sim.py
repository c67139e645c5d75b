"""CPU: the model of the tensor-core samplers' intra-CTA hand-off protocol (tools/tc_protocol_sim.py) finds no deadlock, stale
operand read or tensor-memory overwrite under adversarial random schedules — and does find the early-release race of the single
a_ready barrier the kernel had before, so the model is able to see that class of bug."""
import importlib.util
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("tc_protocol_sim", os.path.join(ROOT, "tools", "tc_protocol_sim.py"))
sim = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sim)


def _violations(runs, steps, old, **kw):
    bad = 0
    for seed in range(runs):
        try:
            sim.Sim(steps, old, random.Random(seed), **kw).run()
        except sim.Violation:
            bad += 1
    return bad


def test_shipped_protocol_has_no_violation():
    # the f16x2 issuer: no accumulator waits (the operand announcements imply the release), fused issue groups
    assert _violations(300, 3, old=False) == 0


def test_explicit_accumulator_waits_have_no_violation():
    # the protocol before that change: one bar_acc_empty wait per unit, one issue group per wait (head units >= 2 of the teams of
    # 2 and 1 still wait this way)
    assert _violations(200, 3, old=False, implied=False) == 0


def test_model_sees_a_false_implied_release():
    # negative control: announcing the second operand half before reading unit b's accumulator breaks the implication
    assert _violations(50, 2, old=False, late_read=True) > 0


def test_model_sees_the_old_a_ready_race():
    assert _violations(50, 2, old=True) > 0


def test_team_exchange_and_grid_word_have_no_violation():
    for seed in range(300):
        sim.TeamSim(6, 3, random.Random(seed)).run()
