from .sequence import DistillSequenceModule, BatchDistillationSequence, TeacherScores, make_score_embeddings  # noqa: F401
from .item import BatchDistillationItem, DistillModule  # noqa: F401,E402
