"""CPU oracle for the DoWnGAN WGAN-GP training iteration.

TEST INFRASTRUCTURE ONLY.  Nothing under ``downgan_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
or the timed CPU baseline — never as the product path.

Parity status: the reference ships no golden vectors or runnable tests for
this path.  The restatement here is pinned instead against
the reference's own importable modules run in the build container
(``oracle/make_golden.py`` imports ``/root/reference/DoWnGAN/networks`` and
asserts bit-exact agreement, then writes ``tests/golden/*.npz``).  The trainer
module of the reference (``DoWnGAN/GAN/wasserstein.py``) cannot be imported
(needs CUDA + xarray + mlflow at import), so the trainer restatement is pinned
by running the reference's *statements* (autograd ``create_graph`` GP) on the
reference's *modules* in ``make_golden.py`` — see that file.
"""
from .networks import (  # noqa: F401
    GeneratorSpec,
    CriticSpec,
    init_generator_state,
    init_critic_state,
    generator_forward,
    critic_forward,
)
from .trainer import (  # noqa: F401
    Hyper,
    gradient_penalty,
    critic_loss_and_grads,
    generator_loss_and_grads,
    gp_param_grads_closed_form,
    AdamState,
    adam_update,
    OracleTrainer,
)
