"""CPU oracle for the rollout hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a checker: a torch-CPU restatement of the reference's task
tensor pipeline and rollout-storage arithmetic, written op-for-op so that fp32 rounding is the
reference's.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product (``massive_marl_benchmark_b200``) never
does: it calls the sm_100a kernels through the C-ABI and fails loudly when the library is
missing.

Pinning status
--------------
* Functions that live in the reference tree (``agents/tasks/*.py``, ``agents/utils/
  torch_jit_utils.py``, ``agents/algorithms/...``): pinned.  ``tests/golden/make_golden.py``
  imports the reference itself (under ``oracle/refshim``) in the build container, checks the
  oracle against it bit for bit and commits the vectors under ``tests/golden``.
* ``isaacgym.torch_utils`` (quat_mul, quat_rotate, get_euler_xyz, ...): the module is a
  proprietary third-party dependency (NVIDIA Isaac Gym Preview 3/4, not pinned in the reference's
  setup.py) and is absent from ``/root/reference``.  ``oracle/isaac_torch_utils.py`` restates the
  published algorithm (public IsaacGymEnvs ``torch_jit_utils.py``).  PARITY UNPINNED at that
  boundary: the reference holds no test or golden vector for it.
"""
