"""Import-time stand-in (bayesopt/utils.py:143-144 imports matplotlib at module level)."""
rcParams = {}
