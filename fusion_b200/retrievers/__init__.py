"""Drop-in mirrors of the reference's ``src/retrievers`` entry points that sit on the scoring hot path."""
