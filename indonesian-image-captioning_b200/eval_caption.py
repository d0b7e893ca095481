"""eval_caption.py -- overlay of the reference's evaluation script.

Same command line and the same `evaluate(args)` entry as the reference's `eval_caption.py`; the per-image loop of
`eval_caption.py:96-131` (DataLoader with batch_size=1, one `decoder.sample` per image, host-side top-k every step)
runs through `capdec.evalcap.generate_captions` -> `decoder.sample_batch` -> `capdec_beam_search_strided`: whole
batches of images per call, the beam search entirely on the device.  Everything around the loop is the reference's:
`datasets.caption.CaptionDataset`, `EncoderCaption`, `EncoderTagger`, `utils.loader.load_decoder` (which resolves the
decoder classes to this overlay), the hypothesis / reference string format (:121-129), the nlg-eval layout (:135-141)
and the three JSON files it writes (:150-165).

    export PYTHONPATH=<this directory>:<reference checkout>
    python <this directory>/eval_caption.py -t attention_scn -mc <checkpoint> -wm <word map> -tm <tag map> -bs 3

What differs from the reference file, on purpose (it cannot run as shipped):
  * `torch` is imported (the reference calls `torch.load` without importing it, :60,70);
  * `CaptionDataset` comes from `datasets.caption` (`datasets/__init__.py` is empty, :15);
  * the tagger runs only for the SCN decoders (:108 calls it unconditionally);
  * the output directory name is `str(current_time)` (:146 joins an int) and the final print uses `str.format` (:189);
  * `--batch_size` (default 32) images go through the beam search together; the reference's `shuffle=True` is kept
    off so that runs are repeatable (the metrics do not depend on the order);
  * an image for which no beam emits `<end>` yields the best live beam and is counted in `incomplete` instead of the
    `ValueError` of `max()` over an empty list (App. C-4);
  * nlg-eval is imported when the scores are computed; where it is not installed the references / hypotheses are
    still written and `evaluate` returns None for the scores.
"""
import argparse
import json
import os
import time

import torch

from capdec import evalcap


def build_parser():
    """The reference's arguments (eval_caption.py:171-185) plus --batch_size / --output_dir."""
    parser = argparse.ArgumentParser(
        description='[(S)how (A)ttend (T)ell - (S)emantic (C)ompositional (N)etworks] - Eval Caption (capdec)')
    parser.add_argument('--type', '-t', help='model type')
    parser.add_argument('--model_caption', '-mc', help='path to pretrained caption model')
    parser.add_argument('--model_tagger', '-mt',
                        default='BEST_checkpoint_tagger_flickr10k_5_cap_per_img_5_min_word_freq.pth.tar',
                        help='path to pretrained tagger model')
    parser.add_argument('--data_folder', '-df', default='./scn_data', help='data folder')
    parser.add_argument('--data_name', '-dn', default='flickr10k_5_cap_per_img_5_min_word_freq', help='data path')
    parser.add_argument('--tag_map', '-tm', help='path to tag map JSON')
    parser.add_argument('--word_map', '-wm', help='path to word map JSON')
    parser.add_argument('--beam_size', '-bs', default=5, type=int, help='beam size')
    parser.add_argument('--batch_size', default=32, type=int, help='images per beam-search call')
    parser.add_argument('--output_dir', default='evaluation', help='where the JSON files go')
    return parser


def _load_models(args, vocab_size, device):
    """eval_caption.py:56-91, with the reference's own modules and loader."""
    from utils.loader import load_decoder, scn_based_model
    need_tag = args.type in scn_based_model
    encoder_tagger = None
    if need_tag:
        from models.encoders.tagger import EncoderTagger
        tagger_checkpoint = torch.load(args.model_tagger, map_location=lambda storage, loc: storage)
        encoder_tagger = EncoderTagger()
        encoder_tagger.load_state_dict(tagger_checkpoint['model_state_dict'])
        encoder_tagger = encoder_tagger.to(device).eval()
    caption_checkpoint = torch.load(args.model_caption, map_location=lambda storage, loc: storage)
    from models.encoders.caption import EncoderCaption
    encoder_caption = EncoderCaption()
    encoder_caption.load_state_dict(caption_checkpoint['encoder_model_state_dict'])
    encoder_caption = encoder_caption.to(device).eval()
    decoder_caption = load_decoder(model_type=args.type, checkpoint=caption_checkpoint['decoder_model_state_dict'],
                                   vocab_size=vocab_size)
    return encoder_caption, encoder_tagger, decoder_caption.eval()


def _make_loader(args):
    """eval_caption.py:38-42; batch_size images at a time."""
    import torchvision.transforms as transforms
    from torch.utils.data import DataLoader
    from datasets.caption import CaptionDataset
    normalize = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    return DataLoader(CaptionDataset(args.data_folder, args.data_name, 'TEST', transform=transforms.Compose([normalize])),
                      batch_size=args.batch_size, shuffle=False, num_workers=1, pin_memory=True)


def evaluate(args, *, loader=None, models=None, word_map=None, metrics=None):
    """Evaluation (eval_caption.py:30-167).  The keyword arguments replace the pieces the reference builds from files
    -- `loader` yields (image, caption, caplen, allcaps) like CaptionDataset('TEST'), `models` is (encoder_caption,
    encoder_tagger, decoder_caption), `metrics(references, hypotheses)` replaces nlg-eval -- so that the host logic is
    testable without a dataset.  Returns the scores dict (None without nlg-eval); the three JSON files are written
    either way, plus `<type>_beam_<k>_incomplete.json` listing images for which no beam ended."""
    if word_map is None:
        with open(args.word_map, 'r') as j:
            word_map = json.load(j)
    if models is None:
        from utils.device import get_device
        models = _load_models(args, len(word_map), get_device())
    encoder_caption, encoder_tagger, decoder_caption = models
    if loader is None:
        loader = _make_loader(args)

    def batches():
        for image, _, _, allcaps in loader:
            yield image, allcaps

    references_temp, hypotheses, completed = evalcap.generate_captions(
        decoder_caption, batches(), word_map, beam_size=args.beam_size, encoder_caption=encoder_caption,
        encoder_tagger=encoder_tagger)
    assert len(references_temp) == len(hypotheses)
    references = evalcap.transpose_references(references_temp)

    out_dir = os.path.join(getattr(args, 'output_dir', 'evaluation'), str(round(time.time())))
    os.makedirs(out_dir, exist_ok=True)
    stem = os.path.join(out_dir, '{}_beam_{}_'.format(args.type, args.beam_size))
    with open(stem + 'references.json', 'w') as f:
        json.dump(references, f)
    with open(stem + 'hypotheses.json', 'w') as f:
        json.dump(hypotheses, f)
    incomplete = [i for i, ok in enumerate(completed) if not ok]
    with open(stem + 'incomplete.json', 'w') as f:
        json.dump(incomplete, f)
    scores = None
    try:
        scores = metrics(references, hypotheses) if metrics is not None else \
            evalcap.compute_metrics(references_temp, hypotheses)
    except ImportError:
        print('nlg-eval is not installed: references / hypotheses written to {}, no scores'.format(out_dir))
    if scores is not None:
        with open(stem + 'scores.json', 'w') as f:
            json.dump(scores, f)
    return scores


if __name__ == '__main__':
    args = build_parser().parse_args()
    score = evaluate(args)
    print("\nScore of {} model @ beam size of {} is {}.\n".format(args.type, args.beam_size, score))
