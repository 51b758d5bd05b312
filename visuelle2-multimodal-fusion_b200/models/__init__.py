"""Drop-in mirrors of the reference's ``models/*.py`` (same class names, constructors, forward
signatures, state_dict keys, Lightning hooks); the bodies call the fused CUDA regions."""
