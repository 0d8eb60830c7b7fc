"""
Apply super resolution to one image file (drop-in for the reference's sr.py).

  python sr.py --file=your_file.png [--scale=3 --layers=8 --filters=96 ...]

Writes the original, bicubic and result images to `<output_dir>/<model name>/` (DCSCN.do_for_file).
Model flags must match the checkpoint, which is looked up as `<checkpoint_dir>/<model name>.ckpt`.
"""

import DCSCN
from helper import args

args.flags.DEFINE_string("file", "image.jpg", "Target filename")
FLAGS = args.get()


def main(_):
    model = DCSCN.SuperResolution(FLAGS, model_name=FLAGS.model_name)
    model.build_graph()
    model.build_optimizer()  # the reference builds it so that its Saver also restores the Adam slots (sr.py:41)
    model.build_summary_saver()
    model.init_all_variables()
    model.load_model()
    model.do_for_file(FLAGS.file, FLAGS.output_dir)


if __name__ == '__main__':
    args.run(main)
