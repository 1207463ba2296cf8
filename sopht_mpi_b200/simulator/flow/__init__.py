from .flow_simulators_mpi_3d import UnboundedFlowSimulator3D
try:  # 2D twin
    from .flow_simulators_mpi_2d import UnboundedFlowSimulator2D
except ImportError:  # pragma: no cover
    pass
