"""openballbot_rl_b200 -- B200-native batched ballbot simulator behind the reference's env / plugin API."""
__version__ = "0.1.0"
