"""Drop-in alias of the reference's `models` package: `importlib.import_module(f"models.{name}.model").TransformerModel`
(reference inference.py:57-58, speed_test.py:34-35) resolves to the B200 engine's classes."""
