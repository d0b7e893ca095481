"""Batched caption generation for evaluation: the per-image loop of the reference's
`eval_caption.py:96-131` (DataLoader with batch_size=1, one `decoder.sample` per image) through
the batched device-side beam search (`sample_batch` -> capdec_beam_search), SURVEY.md §8 f4.

The driver takes batches of ANY size.  The hypothesis / reference strings are formatted exactly as
the reference does (`<start>`, `<end>`, `<pad>` dropped, words joined by one space,
eval_caption.py:121-129), and `transpose_references` is its re-shaping for nlg-eval (:135-141).
The ResNet-152 encoder and the tagger stay the reference's own modules: pass them in, or pass
pre-computed features.  nlg-eval itself is not part of this package (absent from the image);
`compute_metrics` calls it when it is importable.
"""
import torch

START, END, PAD = "<start>", "<end>", "<pad>"


def _strip(seq, drop):
    return [w for w in seq if w not in drop]


def generate_captions(decoder, batches, word_map, beam_size=3, encoder_caption=None, encoder_tagger=None,
                      max_steps=50, device=None):
    """batches: iterable of (images_or_features, allcaps) or (images_or_features, tags, allcaps).
      * with `encoder_caption` (and `encoder_tagger` for the SCN decoders) the first item is the image
        batch (G, 3, H, W), as `CaptionDataset` yields it (eval_caption.py:96-107);
      * without encoders it is the feature batch (G, 14, 14, E) and, for the SCN decoders, `tags`
        (G, S) is the second item.
    allcaps: (G, captions_per_image, L) token ids, or None.
    Returns (references_temp, hypotheses, completed): per image the list of reference strings, the
    hypothesis string, and whether a beam emitted <end> (the reference's `sample` raises where none
    does, App. C-4; here the best live beam is reported and flagged)."""
    rev = {v: k for k, v in word_map.items()}
    drop = {word_map[START], word_map[END], word_map[PAD]}
    need_tag = decoder.kind != "pure_attention"
    dev = device or next(decoder.parameters()).device
    references_temp, hypotheses, completed = [], [], []
    was_training = decoder.training
    decoder.eval()
    try:
        with torch.no_grad():
            for batch in batches:
                if len(batch) == 3:
                    first, tags, allcaps = batch
                else:
                    first, allcaps = batch
                    tags = None
                first = first.to(dev)
                if encoder_caption is not None:
                    enc = encoder_caption(first)
                    if need_tag:
                        if encoder_tagger is None:
                            raise ValueError("encoder_tagger is required for %s" % decoder.kind)
                        tags = encoder_tagger(first)
                else:
                    enc = first
                    if need_tag:
                        if tags is None:
                            raise ValueError("tags are required for %s" % decoder.kind)
                        tags = tags.to(dev)
                res = decoder.sample_batch(beam_size, word_map[START], word_map[END], enc,
                                           tags if need_tag else None, max_steps=max_steps,
                                           want_alphas=False)
                seqs, lens, done = res["seq"].cpu(), res["len"].cpu(), res["completed"].cpu()
                for g in range(seqs.size(0)):
                    seq = seqs[g, :int(lens[g])].tolist()
                    hypotheses.append(" ".join(rev[w] for w in _strip(seq, drop)))
                    completed.append(bool(done[g]))
                    if allcaps is not None:
                        references_temp.append([" ".join(rev[w] for w in _strip(c, drop))
                                                for c in allcaps[g].tolist()])
    finally:
        decoder.train(was_training)
    return references_temp, hypotheses, completed


def transpose_references(references_temp):
    """[image][caption] -> [caption][image], the layout nlg-eval reads (eval_caption.py:135-141)."""
    if not references_temp:
        return []
    refs = [[] for _ in range(len(references_temp[0]))]
    for per_image in references_temp:
        for i, r in enumerate(per_image):
            refs[i].append(r)
    return refs


def compute_metrics(references_temp, hypotheses):
    """BLEU / METEOR / ROUGE / CIDEr through nlg-eval as the reference does (eval_caption.py:147,161);
    raises ImportError where nlg-eval is not installed."""
    from nlgeval import NLGEval
    n = NLGEval(no_skipthoughts=True, no_glove=True)
    return n.compute_metrics(ref_list=transpose_references(references_temp), hyp_list=hypotheses)
