"""chiron_b200 -- B200-native basecalling inference path with the `chiron call` / chiron_eval.evaluation() surface."""
__version__ = "0.1.0"
