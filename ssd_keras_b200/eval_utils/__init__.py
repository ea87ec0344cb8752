"""Evaluation on the device: the matching of predictions to ground truth that the reference's
``Evaluator`` does in a per-prediction Python loop."""
