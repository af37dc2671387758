"""ORACLE — test infrastructure only (see oracle/ops.py header). Never imported by the product."""
