from .verification import extract_embeddings, random_block_occlusion, test  # noqa: F401
