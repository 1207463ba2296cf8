from .immersed_body_forcing_grid import (EmptyForcingGrid, ImmersedBodyForcingGrid,
                                         PrescribedForcingGrid)
from .immersed_body_flow_interaction_mpi import (CosseratRodFlowInteraction,
                                                 ImmersedBodyFlowInteractionMPI,
                                                 RigidBodyFlowInteractionMPI)
from .forcing_grids import (CircularCylinderForcingGrid, CosseratRodElementCentricForcingGrid,
                            CosseratRodSurfaceForcingGrid, FlowForces, SphereForcingGrid)
