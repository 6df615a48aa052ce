"""RL: the package name the reference imports itself under (`from RL.Chief import Chief`,
/root/reference/main.py:8-9, PPO.py:4-5).  Same module and class names, same constructor and
method signatures; TensorFlow graph/session objects are replaced by thin handles that dispatch
to the sm_100a engine (dppo_b200.DppoEngine).  No TensorFlow, no CPU compute path."""
