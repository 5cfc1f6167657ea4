"""The linearised partial section of the lane-cooperative Poseidon-Goldilocks permutation (csrc/poseidon_g_coop2.cuh):
the committed tables are what tools/gen_poseidon_coop2_constants.py derives, and the schedule the kernel runs on them
equals the fast form of chip/plonk/gates/poseidon.rs:634-686 in big-integer arithmetic (known answer, corner states,
random states, and the independent pure-Python permutation of tests/pyref)."""
import importlib.util
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_coop2", os.path.join(ROOT, "tools", "gen_poseidon_coop2_constants.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_schedule_equals_fast_form_and_pyref():
    g = _gen()
    B, F = g.derive()
    ZC, QC, C0 = g.lane_tables(B, F)
    assert g.permute_fast(list(range(12)))[0] == 0xd64e1e3efc5b8e9e
    from pyref import poseidon as pp
    rnd = random.Random(11)
    P = g.P
    cases = [list(range(12)), [0] * 12, [P - 1] * 12, [0xFFFFFFFF] * 12, [0xFFFFFFFF00000000] * 12]
    cases += [[rnd.randrange(P) for _ in range(12)] for _ in range(25)]
    perm = pp.permute_g
    for s in cases:
        want = g.permute_fast(s)
        assert g.permute_coop2(s, ZC, QC, C0) == want
        assert [int(x) for x in perm(list(s))] == want


def test_rows_have_the_shape_the_kernel_assumes():
    g = _gen()
    B, F = g.derive()
    ZC, QC, C0 = g.lane_tables(B, F)
    # row 0 of QC is the dummy "p_{-1}" row; B'_r may only use p_0 .. p_{r-2} (it is broadcast in round r-1, one round behind the chain)
    assert all(v == 0 for s in range(3) for v in QC[0][s])
    for r in range(22):
        slot, lane = divmod(r, 16)
        for j in range(1, 23):            # QC[j] multiplies p_{j-1}
            if j - 1 >= max(r - 1, 0):
                assert QC[j][slot][lane] == 0, (r, j)
    # unused lanes carry zero rows (lanes 12..15 of the F slot, lanes 6..15 of slot 1)
    for l in range(12, 16):
        assert C0[2][l] == 0 and all(ZC[k][2][l] == 0 for k in range(12)) and all(QC[j][2][l] == 0 for j in range(23))
    for l in range(6, 16):
        assert C0[1][l] == 0 and all(ZC[k][1][l] == 0 for k in range(12)) and all(QC[j][1][l] == 0 for j in range(23))


def test_committed_tables_are_current(tmp_path):
    g = _gen()
    committed = open(g.OUT).read()
    g.OUT = tmp_path / "coop2.inc"
    g.main()
    assert open(g.OUT).read() == committed
