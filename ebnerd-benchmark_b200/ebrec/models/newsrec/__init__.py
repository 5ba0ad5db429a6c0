"""B200-native drop-in for ``ebrec.models.newsrec`` (reference: src/ebrec/models/newsrec/__init__.py:1-4)."""
