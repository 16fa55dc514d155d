"""Test-infrastructure oracles (CPU restatements of the reference path). Never imported by the product."""
