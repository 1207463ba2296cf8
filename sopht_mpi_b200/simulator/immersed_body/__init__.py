from .immersed_body_forcing_grid import (EmptyForcingGrid, ImmersedBodyForcingGrid,
                                         PrescribedForcingGrid)
from .immersed_body_flow_interaction_mpi import (CosseratRodFlowInteraction,
                                                 ImmersedBodyFlowInteractionMPI,
                                                 RigidBodyFlowInteractionMPI)
