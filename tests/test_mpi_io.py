"""Checkpoint / restart through the reference's MPIIO registration API (npz container)."""
import numpy as np
import pytest
import torch

from sopht_mpi_b200.utils import MPIIO, DeviceField, MPIConstruct2D, MPIConstruct3D


@pytest.mark.parametrize("dim", [2, 3])
def test_save_and_load_round_trip(tmp_path, dim):
    real_t, gs = np.float64, 2
    n = (6, 8, 10)[3 - dim:]
    mc = (MPIConstruct3D if dim == 3 else MPIConstruct2D)(*n, real_t=real_t)
    rng = np.random.default_rng(0)
    shape = tuple(v + 2 * gs for v in n)
    vort = rng.uniform(size=(dim,) + shape)
    scal = DeviceField(torch.from_numpy(rng.uniform(size=shape)))  # a "device" field (CPU tensor here)
    pos = rng.uniform(size=(dim, 7))
    mismatch = rng.uniform(size=(dim, 7))
    lag_scalar = rng.uniform(size=7)

    def make_io(v, s, p, m, ls):
        io = MPIIO(mpi_construct=mc, real_dtype=real_t)
        io.define_eulerian_grid(origin=np.zeros(dim), dx=np.full(dim, 0.1), grid_size=np.array(n), ghost_size=gs)
        io.add_as_eulerian_fields_for_io(vorticity=v, scalar=s)
        io.add_as_lagrangian_fields_for_io(lagrangian_grid=p, lagrangian_grid_master_rank=0,
                                           lagrangian_grid_name="mismatch_field", position_mismatch=m, tag=ls)
        return io

    io = make_io(vort, scal, pos, mismatch, lag_scalar)
    assert io.eulerian_fields_type == {"vorticity": "Vector", "scalar": "Scalar"}
    assert io.lagrangian_fields_type == {"position_mismatch": "Vector", "tag": "Scalar"}
    path = str(tmp_path / "flow_00010.h5")
    io.save(h5_file_name=path, time=0.375)
    inner = (slice(gs, -gs),) * dim
    v2 = np.full_like(vort, 7.0)
    s2 = DeviceField(torch.full(shape, 7.0, dtype=torch.float64))
    p2, m2, l2 = np.zeros_like(pos), np.zeros_like(mismatch), np.zeros_like(lag_scalar)
    t = make_io(v2, s2, p2, m2, l2).load(h5_file_name=path)
    assert t == 0.375
    assert np.array_equal(v2[(slice(None),) + inner], vort[(slice(None),) + inner])
    assert np.all(v2[:, 0] == 7.0)  # ghost cells are not part of a checkpoint
    assert np.array_equal(np.asarray(s2)[inner], np.asarray(scal)[inner])
    assert np.array_equal(p2, pos) and np.array_equal(m2, mismatch) and np.array_equal(l2, lag_scalar)
    with pytest.raises(ValueError):
        bad = MPIIO(mpi_construct=mc, real_dtype=real_t)
        bad.define_eulerian_grid(origin=np.zeros(dim), dx=np.full(dim, 0.2), grid_size=np.array(n), ghost_size=gs)
        bad.add_as_eulerian_fields_for_io(vorticity=v2)
        bad.load(path)
    with pytest.raises(ValueError):
        io.add_as_eulerian_fields_for_io(wrong=np.zeros((3, 3)))
