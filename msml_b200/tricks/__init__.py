from .consensus_loss import StructureConsensuLossFunction  # noqa: F401
