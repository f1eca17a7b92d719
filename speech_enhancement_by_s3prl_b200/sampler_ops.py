"""Active-sampling scoring (reference sampler.py:59-120, called from the sampler child process that
runner.py:383-411 drives): per-utterance gradient embeddings of the mask head and their cosine match
against the mean query gradient.

``scoring`` in the reference runs one ``backward(retain_graph=True)`` per utterance.  For the
projection heads of the named path (``Linear`` / ``LinearResidual`` with the spectral SISDR
objective) every utterance's gradient comes out of ONE pass here: the per-utterance loss terms do
not interact, so d loss_u / d offset for the whole batch is one launch of the objective's backward
with grad_out = 1, and the split-K weight-gradient kernel run with one split per utterance leaves
exactly the per-utterance (grad_W, grad_b) in its partials (``se_head_grad_embeddings``).
Other heads / objectives (the recurrent ones are cuDNN, outside the path) go through
``scoring_loop`` -- the reference's loop on the drop-in modules.
"""
import torch

from . import _lib, model, objective, ops

get_length_masks = ops.length_masks                                   # sampler.py:35-39


def matching(query_scores, key_scores, eps=1e-12):
    """sampler.py:113-116 in one pass over each matrix: (n_query, P), (n_key, P) -> (n_key,)."""
    q, k = ops._c(query_scores, "query_scores"), ops._c(key_scores, "key_scores")
    if q.dim() != 2 or k.dim() != 2 or q.shape[1] != k.shape[1]:
        raise RuntimeError(f"matching: shapes {tuple(q.shape)} and {tuple(k.shape)} do not agree")
    nq, P = q.shape
    nk = k.shape[0]
    with torch.cuda.device(q.device):
        ws_d = torch.empty(nq + 2 * nk, device=q.device, dtype=torch.float64)
        ws_q = torch.empty(P, device=q.device)
        out = torch.empty(nk, device=q.device)
        rc = _lib.load().se_match_scores(q.data_ptr(), nq, k.data_ptr(), nk, P, float(eps), ws_d.data_ptr(), ws_q.data_ptr(),
                                         out.data_ptr(), ops._stream())
        _lib.check(rc, "se_match_scores")
    return out


def thresholding(match_scores):
    """sampler.py:119-120."""
    return match_scores > 0


def head_grad_embeddings(features, offset, grad_offset, activation, mean=None, std=None, cmvn_eps=1e-6):
    """(B, F, Din) features, (B, F, Dout) offset and d loss_u / d offset -> (B, Dout*Din + Dout) rows
    [grad_W.view(-1), grad_b] per utterance, in ``named_parameters`` order (sampler.py:95-108)."""
    x, off, go = ops._c(features, "features"), ops._c(offset, "offset"), ops._c(grad_offset, "grad_offset")
    B, F, Din = x.shape
    Dout = off.shape[2]
    lib = _lib.load()
    with torch.cuda.device(x.device):
        ws_floats = lib.se_head_grad_embeddings_workspace(B, F, Din, Dout)
        if ws_floats <= 0:
            raise RuntimeError(f"se_head_grad_embeddings: shape (B={B}, F={F}, D_in={Din}, D_out={Dout}) is not supported")
        ws = torch.empty(ws_floats, device=x.device)
        out = torch.empty(B, Dout * Din + Dout, device=x.device)
        rc = lib.se_head_grad_embeddings(x.data_ptr(), Din, ops._p(mean), ops._p(std), None, Din, float(cmvn_eps), off.data_ptr(),
                                         go.data_ptr(), Dout, B, F, Din, Dout, ops.ACT[activation], ws.data_ptr(), ws_floats,
                                         out.data_ptr(), ops._stream())
        _lib.check(rc, "se_head_grad_embeddings")
    return out


def _batched_ok(head, criterion, features):
    if type(head) not in (model.Linear, model.LinearResidual) or type(criterion) is not objective.SISDR:
        return False
    B, F, Din = features.shape
    return _lib.load().se_head_grad_embeddings_workspace(B, F, Din, head.linear.weight.shape[0]) > 0


def _projection_of(head):
    """(nn.Linear, activation, input transform) of a head whose LAST layer emits the log-spectrum the L1 objective reads
    (objective.py:109-117): ``LSTM`` (model.py:57-60: log_predicted = scaling_layer(lstm(x))) or a bare ``Linear``."""
    if type(head) is model.LSTM:
        return head.scaling_layer[0], head.activation, (lambda x: head.lstm(x)[0])
    if type(head) is model.Linear:
        return head.linear, head.activation, (lambda x: x)
    return None


def _batched_l1_ok(head, criterion, features, projection_only):
    if type(criterion) is not objective.L1:
        return False
    proj = _projection_of(head)
    if proj is None or (type(head) is model.LSTM and not projection_only):
        return False                                     # gradients of the recurrent layers need one BPTT per utterance
    B, F, _ = features.shape
    return _lib.load().se_head_grad_embeddings_workspace(B, F, proj[0].weight.shape[1], proj[0].weight.shape[0]) > 0


@torch.no_grad()
def scoring_batched_l1(head, criterion, features, linear_tar, stft_lengths, mean=False):
    """Gradient embeddings of the projection layer under objective.L1 (sampler.py:72-110 with the objective run_active.sh
    names) without a Python loop: log_predicted = act(W h + b) from the head kernel, d L1 / d log_predicted = sign(log_predicted -
    log(linear_tar + eps)) on the valid frames from the objective's backward kernel (ONE launch for the batch), and the
    split-K weight-gradient kernel with one split per utterance.  Row u is the gradient of the L1 of utterance u ALONE (a
    mean over ITS valid elements, what the reference's loop computes); mean=True: gradient of the L1 of the whole batch
    (a mean over ALL valid elements).  Rows are [grad_W.view(-1), grad_b] of the projection."""
    lin, act, body = _projection_of(head)
    x = ops._c(body(features), "features")
    B, F, _ = x.shape
    K = lin.weight.shape[0]
    log_pred = ops.linear_head(x, lin.weight, lin.bias, act, precision=head.precision)
    tar = ops._c(linear_tar, "linear_tar")
    frames = ops._c(stft_lengths, "stft_lengths", torch.int64)
    sign = torch.ops.se_b200.l1_logspec_bwd(log_pred, tar, frames, float(criterion.eps), 1.0, torch.ones(1, device=x.device))
    grads = head_grad_embeddings(x, log_pred, sign, act)                      # per-utterance SUMS of sign * d log_pred / d theta
    count = (frames.clamp(max=F) * K).to(grads.dtype)
    if mean:
        return grads.sum(dim=0, keepdim=True) / count.sum()
    return grads / count.unsqueeze(1)


@torch.no_grad()
def scoring_batched(head, criterion, features, linear_inp, linear_tar, stft_lengths, mean=False):
    """Gradient embeddings of a Linear / LinearResidual head under objective.SISDR without a Python loop.
    mean=False: (B, P), row u = gradient of the loss of utterance u alone (what criterion returns for a batch of one);
    mean=True: (1, P), gradient of the batch-mean loss."""
    x = ops._c(features, "features")
    B, F, Din = x.shape
    K = head.linear.weight.shape[0]
    mean_t = std_t = None
    if isinstance(head, model.LinearResidual):
        if head.cmvn:
            mean_t, std_t = ops.cmvn_stats(x)
        offset = ops.linear_head(x, head.linear.weight, head.linear.bias, head.activation, mean_t, std_t, head.eps,
                                 precision=head.precision)
        lin = ops._c(linear_inp, "linear_inp")
        eps_c = head.eps
    else:                                                             # Linear: predicted = act(W x + b), no mask multiply
        offset = ops.linear_head(x, head.linear.weight, head.linear.bias, head.activation, precision=head.precision)
        lin = None
        eps_c = 1e-6
    tar = ops._c(linear_tar, "linear_tar")
    frames = ops._c(stft_lengths, "stft_lengths", torch.int64)
    if lin is not None:
        _, sums3 = ops.sisdr_mask_fwd(offset, lin, tar, frames, K, criterion.eps)
        go = ops.sisdr_mask_bwd(offset, lin, tar, frames, K, sums3, torch.ones(B, device=x.device), criterion.eps)
    else:
        _, sums3 = ops.sisdr_mask_fwd(None, offset, tar, frames, K, criterion.eps)
        go = ops.sisdr_mask_bwd(None, offset, tar, frames, K, sums3, torch.ones(B, device=x.device), criterion.eps)
    grads = head_grad_embeddings(x, offset, go, head.activation, mean_t, std_t, eps_c)
    return grads.mean(dim=0, keepdim=True) if mean else grads


def scoring_loop(head, criterion, features, linear_inp, linear_tar, stft_lengths, mean=False, projection_only=False):
    """The reference's loop (sampler.py:77-110) on the drop-in modules: one backward per utterance."""
    predicted, results = head(features=features, linears=linear_inp)
    extra = {k: v for k, v in results.items()}
    if type(head) is model.Linear and "log_predicted" not in extra:
        extra["log_predicted"] = predicted               # a bare projection used as the log-spectrum predictor (L1)
    proj = _projection_of(head) if projection_only else None
    keep = None if proj is None else {id(p) for p in proj[0].parameters()}
    idx = [slice(None)] if mean else [slice(u, u + 1) for u in range(predicted.shape[0])]
    grads = []
    for sl in idx:
        kw = {k: v[sl] for k, v in extra.items()}
        loss, _ = criterion(predicted=predicted[sl], linear_tar=linear_tar[sl], linear_inp=linear_inp[sl],
                            stft_lengths=stft_lengths[sl], **kw)
        head.zero_grad()
        loss.backward(retain_graph=True)
        grads.append(torch.cat([p.grad.reshape(-1) for _, p in head.named_parameters()
                                if p.grad is not None and (keep is None or id(p) in keep)]).detach())
        head.zero_grad()
    return torch.stack(grads, dim=0)


def _fused_ok(preprocessor, head, criterion, B, T, feat_log):
    """The whole scoring pass on the engine's row-padded tensors (register-resident STFT with fused log / CMVN sums, TMA head, the
    objective's backward folded into the per-utterance weight-gradient kernel): LinearResidual on log-power + SISDR, tensor-core head."""
    if type(head) is not model.LinearResidual or type(criterion) is not objective.SISDR or not feat_log or getattr(head, "precision", 0) != 1:
        return False
    n_fft, hop = preprocessor._win_args["n_fft"], preprocessor._win_args["hop_length"]
    K, F = n_fft // 2 + 1, T // hop + 1
    LD = ops.round4(K)
    if tuple(head.linear.weight.shape) != (K, K):
        return False
    lib = _lib.load()
    return bool(lib.se_linear_head_fused_supported(B, F, K, K, LD, LD, LD)) and \
        bool(lib.se_head_grad_embeddings_sisdr_supported(B, F, K, K, LD, LD, LD, LD))


@torch.no_grad()
def scoring_fused(preprocessor, head, criterion, lengths, wavs, mean=False):
    """sampler.py:59-110 in seven launches: K1 (noisy: power + log-power + CMVN sums), K1 (clean: power), TMA head, SISDR sums +
    finish, per-utterance weight gradient with the objective's backward folded in, pack.  Rows as ``scoring_batched``."""
    B, C, T = wavs.shape
    dev = wavs.device
    n_fft, hop = preprocessor._win_args["n_fft"], preprocessor._win_args["hop_length"]
    K = n_fft // 2 + 1
    LD = ops.round4(K)
    window = preprocessor._frame_window
    if window.device != dev:
        preprocessor.to(dev)
        window = preprocessor._frame_window
    ch_i, ch_t = int(getattr(preprocessor, "channel_inp", 0)), int(getattr(preprocessor, "channel_tar", 1))
    ws = torch.zeros(B * (2 * LD + 3), device=dev, dtype=torch.float64)
    stat_sums, sums3 = ws[:B * 2 * LD].view(B, LD, 2), ws[B * 2 * LD:].view(B, 3)
    linear_inp, logp, linear_tar, _ = ops.stft_features_pair(wavs, ch_i, ch_t, n_fft, hop, window, log_eps=preprocessor.eps,
                                                             stat_sums=stat_sums)
    wpad = ops.round_tf32(ops.pad_weight(head.linear.weight.detach()))
    stats = stat_sums if head.cmvn else None
    offset = ops.linear_head_tma(logp, K, wpad, head.linear.bias, head.activation, stats, head.eps)
    lens = lengths.to(device=dev, dtype=torch.int64).contiguous()
    ops.sisdr_mask_step(offset, linear_inp, linear_tar, lens, hop, K, criterion.eps, sums3=sums3, sums_zeroed=True, want_grad=False)
    F = logp.shape[1]
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws_floats = lib.se_head_grad_embeddings_workspace(B, F, K, K)
        wsf = torch.empty(ws_floats, device=dev)
        out = torch.empty(B, K * K + K, device=dev)
        rc = lib.se_head_grad_embeddings_sisdr(logp.data_ptr(), LD, ops._p(stats), LD, float(head.eps), offset.data_ptr(), offset.shape[2],
                                               linear_inp.data_ptr(), linear_inp.shape[2], linear_tar.data_ptr(), linear_tar.shape[2],
                                               lens.data_ptr(), int(hop), sums3.data_ptr(), float(criterion.eps), B, F, K, K,
                                               ops.ACT[head.activation], wsf.data_ptr(), ws_floats, out.data_ptr(), ops._stream())
        _lib.check(rc, "se_head_grad_embeddings_sisdr")
    return out.mean(dim=0, keepdim=True) if mean else out


def scoring(preprocessor, head, criterion, lengths, wavs, mean=False, feat_log=True, projection_only=False):
    """sampler.py:59-110 for ``--from_rawfeature``: (B, 3, T) batch -> (B, n_params) gradient embeddings
    ((1, n_params) with mean=True).  projection_only: score with the gradients of the head's last (projection) layer only --
    for the recurrent ``LSTM`` head that is the part the library computes in one pass; all parameters go through the loop."""
    if wavs.is_cuda and _fused_ok(preprocessor, head, criterion, wavs.shape[0], wavs.shape[2], feat_log):
        return scoring_fused(preprocessor, head, criterion, lengths, wavs, mean=mean)
    c = preprocessor.get_feat_config
    ch_i, ch_t = int(getattr(preprocessor, "channel_inp", 0)), int(getattr(preprocessor, "channel_tar", 1))
    feats, linear_inp, linear_tar = preprocessor(wavs, [c("linear", ch_i, log=feat_log), c("linear", ch_i), c("linear", ch_t)])
    frames = lengths.to(feats.device) // preprocessor._win_args["hop_length"] + 1
    if _batched_ok(head, criterion, feats):
        return scoring_batched(head, criterion, feats, linear_inp, linear_tar, frames, mean=mean)
    if _batched_l1_ok(head, criterion, feats, projection_only):
        return scoring_batched_l1(head, criterion, feats, linear_tar, frames, mean=mean)
    return scoring_loop(head, criterion, feats, linear_inp, linear_tar, frames, mean=mean, projection_only=projection_only)
