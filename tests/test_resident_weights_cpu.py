"""CPU suite: the bookkeeping that decides when a weight matrix's staged copy may stay resident (Executor._track_weights,
nodes_blas.staged_weight / forget_weights).  A resident copy is only sound if its key can never be presented by different
content: keys are process-unique serials (not id()s), move with the tensor's torch version, and every copy staged for a
tensor is dropped when the tensor is replaced, changes version, or its Executor dies."""

import gc

from pytensor_b200.vm import nodes_blas
from pytensor_b200.vm.nodes_blas import Dot22Node
from pytensor_b200.vm.vm import Executor, Program, Step


class FakeCudaTensor:
    is_cuda = True

    def __init__(self):
        self._version = 0


def _executor():
    prog = Program(3, [0, 1], [2], {}, [Step(Dot22Node("float32"), [0, 1], [2])])
    return Executor(prog, allow_gc=True, use_graph=False)


def test_keys_appear_on_the_third_identical_call_and_are_unforgeable():
    ex = _executor()
    assert ex._w_in == {1: 1}                       # input position 1 is the B operand of the Dot22 step
    x, w = FakeCudaTensor(), FakeCudaTensor()
    assert ex._track_weights([x, w]) == ()          # first sight
    assert ex._track_weights([x, w]) == ()          # same object, same version: once
    k3 = ex._track_weights([x, w])                  # twice in a row -> stable
    assert len(k3) == 1 and k3[0][0] == 1 and k3[0][1][0] == "in"
    assert ex._track_weights([x, w]) == k3          # the key is part of the CUDA-graph signature: it must not move
    # another object — even one that could inherit the id of a dead tensor — starts over with a NEW serial
    w2 = FakeCudaTensor()
    assert ex._track_weights([x, w2]) == ()
    ex._track_weights([x, w2])
    k_new = ex._track_weights([x, w2])
    assert k_new and k_new[0][1][1] != k3[0][1][1]
    # an in-place change (torch bumps the version) makes the tensor volatile for good: staging goes back into the graph
    w2._version += 1
    for _ in range(5):
        assert ex._track_weights([x, w2]) == ()
    # host arrays are never resident
    import numpy as np

    assert ex._track_weights([x, np.zeros((2, 2), dtype="float32")]) == ()


def test_staged_copies_are_dropped_with_their_tensor_version_and_executor(monkeypatch):
    dropped = []
    monkeypatch.setattr(nodes_blas, "forget_weights", lambda kind, serial: dropped.append((kind, serial)))
    ex = _executor()
    x, w = FakeCudaTensor(), FakeCudaTensor()
    for _ in range(3):
        keys = ex._track_weights([x, w])
    serial = keys[0][1][1]
    w._version += 1
    ex._track_weights([x, w])
    assert ("in", serial) in dropped                 # version moved: the old content's copies are gone
    w3 = FakeCudaTensor()
    ex._track_weights([x, w3])
    assert dropped.count(("in", serial)) >= 2        # replaced: dropped again (idempotent)
    uid = ex._uid
    new_serial = ex._w_track[1][4]
    del ex
    gc.collect()
    assert ("in", new_serial) in dropped and ("const", uid) in dropped


def test_forget_weights_removes_only_the_matching_keys():
    class S:
        def __init__(self, n):
            class B:
                def numel(self_inner):
                    return n
            self.buf, self.in_graph = B(), True

    nodes_blas._stage_cache.clear()
    nodes_blas._stage_cache_stats["bytes"] = 0
    for key, n in ((("in", 7, 0), 10), (("in", 7, 1), 20), (("in", 8, 0), 30), (("const", 7, 3), 40)):
        nodes_blas._stage_cache[(key, 3, (4, 4), (4, 1))] = S(n)
        nodes_blas._stage_cache_stats["bytes"] += n
    nodes_blas.forget_weights("in", 7)
    assert sorted(k[0] for k in nodes_blas._stage_cache) == [("const", 7, 3), ("in", 8, 0)]
    assert nodes_blas._stage_cache_stats["bytes"] == 70
    nodes_blas._stage_cache.clear()
    nodes_blas._stage_cache_stats["bytes"] = 0
