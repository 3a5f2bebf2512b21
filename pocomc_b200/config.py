"""Run-time switches that have no counterpart in the reference.

rng_mode
    ``"host"`` (default): every random number of the MCMC step is drawn on the host from the
    legacy global ``np.random`` stream in the reference's order and uploaded --
    bit-faithful to the reference for a given ``random_state``.
    ``"device"``: Philox4x32-10 counters on the GPU keyed by (seed, step, global particle id);
    statistically equivalent, no per-step noise upload, independent of the number of GPUs.
mean_mode
    1 (default with rng_mode "host") reproduces ``np.mean(theta_f32, axis=0)``'s sequential f32
    accumulation exactly (mcmc.py:156); 0 uses deterministic f64 block partial sums.
device_callbacks
    False (default): prior and likelihood are host black boxes like the reference's (x' crosses
    PCIe every MCMC step).  True: a likelihood object exposing ``device(x, finite, out)`` and a
    ``Prior`` of frozen scipy ``norm`` / ``uniform`` factors are evaluated on the GPU instead.
device_prior
    True (default): when the prior handed to the MCMC kernels is ``pocomc_b200.Prior.logpdf`` of
    frozen scipy ``norm`` / ``uniform`` factors, log-prior values are computed on the GPU
    (identical formula, f64); any other prior object is called on the host like the reference does.
inverse_path
    ``"tri"`` (default): ``Flow.inverse`` (and the flow pull-back of every MCMC step) of affine flows whose
    accumulators fit tensor memory (``tri_layout.tri_supported``: D >= 8 at the preset widths) runs on the tcgen05
    block-triangular sweep (csrc/flow_tri.cu, 3xTF32 split = fp32 fidelity); ``"sweep"``: always the fp32-FMA sweep.
forward_path
    ``"tc"`` (default): ``Flow.forward`` / ``log_prob`` without a graph run on the tcgen05 dense kernel
    (csrc/flow_tc.cu, 3xTF32 split = fp32 fidelity) when the flow is affine with H <= 128 and the batch
    has at least ``tc_min_rows`` rows, and on the block-triangular tcgen05 sweep (csrc/flow_tri.cu) when it is affine and
    wider (H = 256 / 512 / 1024); ``"sweep"``: always the degree-ordered fp32-FMA sweep kernel.
fit_path
    ``"graph"`` (default): every optimiser step of ``Flow.fit`` is one CUDA-graph launch (batch gather,
    autograd forward/backward, fused clip + AdamW kernel, loss accumulation) -- same arithmetic and RNG
    consumption as ``"eager"``, which issues the ops one by one.  Noise / L1 / L2 regularised fits
    always take the eager path.
fit_kernels
    ``"fused"`` (default): MAF flows with H <= 256, D <= 64 train on the fused forward/backward kernels of
    csrc/flow_train.cu (4 launches per optimiser step inside the graph), every other flow -- spline heads (the reference's
    default presets), H >= 512 -- on the layer-wise kernels of csrc/flow_train_lw.cu; ``"autograd"``: torch autograd
    over the flat blob (what noise / L1 / L2 regularised fits always use).
p2p_exchange
    True (default): under torch.distributed/NCCL with one GPU per rank the per-MCMC-step reduction of a particle-sharded
    run (mean acceptance, mean theta, tracked log-density over ALL particles) is exchanged by the accept kernel itself
    through peer memory over NVLink (csrc/mcmc_ops.cu: mh_accept_kernel<2>, pmc_comm_*); False: NCCL all-gather of the block
    partials followed by pmc_mcmc_finalize (also what gloo test set-ups with two ranks on one GPU use).
host_chunks
    0 (default): automatic.  With a host likelihood, x' leaves the GPU in this many contiguous row chunks per MCMC step and
    the likelihood is called once per chunk as soon as the chunk has landed, so the remaining copies overlap the host's
    work (4 chunks once x' is 16 MB or more, otherwise 1; always 1 with blobs).  Row-wise independence of the
    likelihood is the reference's own contract (``vectorize`` / ``pool``, sampler.py:807-861).  1: one call per step.
fuse_prior
    True (default): the device form of the prior is evaluated inside the reparameterisation kernel
    (``pmc_scaler_inverse_prior``, one launch less per step, same numbers); False: ``pmc_scaler_inverse`` + ``pmc_logprior``.
"""
import os

rng_mode = os.environ.get("PMC_B200_RNG", "host")
mean_mode = None  # None -> 1 for "host", 0 for "device"
fit_kernels = os.environ.get("PMC_B200_FIT_KERNELS", "fused")
fit_path = os.environ.get("PMC_B200_FIT", "graph")
forward_path = os.environ.get("PMC_B200_FORWARD", "tc")
inverse_path = os.environ.get("PMC_B200_INVERSE", "tri")
tri_passes = int(os.environ.get("PMC_B200_TRI_PASSES", "3"))
tc_min_rows = int(os.environ.get("PMC_B200_TC_MIN_ROWS", "1"))
device_prior = os.environ.get("PMC_B200_DEVICE_PRIOR", "1") == "1"
device_callbacks = os.environ.get("PMC_B200_DEVICE_CALLBACKS", "0") == "1"
# experimental: under torch.distributed every rank stores only its block of the particle history (pocomc_b200.sharded)
shard_history = os.environ.get("PMC_B200_SHARD_HISTORY", "0") == "1"
p2p_exchange = os.environ.get("PMC_B200_P2P", "1") == "1"
tri_rqs_min_hidden = int(os.environ.get("PMC_B200_TRI_RQS_MIN_H", "0"))
host_chunks = int(os.environ.get("PMC_B200_HOST_CHUNKS", "0"))
fuse_prior = os.environ.get("PMC_B200_FUSE_PRIOR", "1") == "1"


def set_rng_mode(mode: str):
    global rng_mode
    if mode not in ("host", "device"):
        raise ValueError("rng mode must be 'host' or 'device'")
    rng_mode = mode


def resolved_mean_mode() -> int:
    if mean_mode is not None:
        return int(mean_mode)
    return 1 if rng_mode == "host" else 0
