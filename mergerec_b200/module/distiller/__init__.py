from .sequence import DistillSequenceModule, BatchDistillationSequence, TeacherScores, make_score_embeddings  # noqa: F401
