"""
Train DCSCN (drop-in for the reference's train.py): same flags, same epoch / learning-rate schedule
(start --initial_lr, x --lr_decay every --lr_decay_epoch epochs, stop below --end_lr), evaluation on
--test_dataset after every epoch, checkpoint (TF V2 bundle) after every epoch and at the end.

  python train.py --dataset=bsd200 --training_images=80000
"""

import logging
import sys

import DCSCN
from helper import args, utilty as util

FLAGS = args.get()


def evaluate_model(model, test_data):
    files = util.get_files_in_directory(FLAGS.data_dir + "/" + test_data)
    scores = [model.do_for_evaluate_with_output(f, output_directory=FLAGS.output_dir, print_console=False)
              for f in files]
    logging.info("Model Average [%s] PSNR:%f, SSIM:%f" % (
        test_data, sum(s[0] for s in scores) / len(files), sum(s[1] for s in scores) / len(files)))


def train(model, flags, trial):
    test_filenames = util.get_files_in_directory(flags.data_dir + "/" + flags.test_dataset)
    if len(test_filenames) <= 0:
        print("Can't load images from [%s]" % (flags.data_dir + "/" + flags.test_dataset))
        exit()

    model.init_all_variables()
    if flags.load_model_name != "":
        model.load_model(flags.load_model_name, output_log=True, restore_optimizer=True)
    model.broadcast_variables()   # data-parallel ranks start from rank 0's weights (no-op for one process)
    model.init_train_step()
    model.init_epoch_index()
    model_updated = True

    psnr, ssim = model.evaluate(test_filenames)
    model.print_status(psnr, ssim, log=True)
    model.log_to_tensorboard(test_filenames[0], psnr, save_meta_data=True)

    while model.lr > flags.end_lr:
        model.build_input_batch()
        model.train_batch()
        if model.training_step * model.batch_num < model.training_images:
            continue
        # one training epoch finished
        model.epochs_completed += 1
        psnr, ssim = model.evaluate(test_filenames)
        model.print_status(psnr, ssim, log=model_updated)
        model.log_to_tensorboard(test_filenames[0], psnr, save_meta_data=model_updated)
        model.save_model(trial=trial, output_log=False)
        model_updated = model.update_epoch_and_lr()
        model.init_epoch_index()

    model.end_train_step()
    model.save_model(trial=trial, output_log=True)  # save last generation anyway

    evaluate_model(model, flags.test_dataset)
    if FLAGS.do_benchmark:
        for test_data in ['set5', 'set14', 'bsd100']:
            if test_data != flags.test_dataset:
                evaluate_model(model, test_data)
    return psnr, ssim


def main(not_parsed_args):
    if len(not_parsed_args) > 1:
        print("Unknown args:%s" % not_parsed_args)
        exit()

    # `torchrun --nproc-per-node N train.py ...`: one process per GPU, mini-batch split over the ranks, one gradient
    # all-reduce per step (NCCL); a plain `python train.py` stays single-process
    rank, world = DCSCN.init_distributed(FLAGS)
    if world > 1:
        print("data-parallel rank %d of %d on GPU %d" % (rank, world, FLAGS.gpu_device_id))
    model = DCSCN.SuperResolution(FLAGS, model_name=FLAGS.model_name)
    if FLAGS.build_batch:
        model.load_datasets(FLAGS.data_dir + "/" + FLAGS.dataset, FLAGS.batch_dir + "/" + FLAGS.dataset,
                            FLAGS.batch_image_size, FLAGS.stride_size)
    else:
        model.load_dynamic_datasets(FLAGS.data_dir + "/" + FLAGS.dataset, FLAGS.batch_image_size)
    model.build_graph()
    model.build_optimizer()
    model.build_summary_saver()

    logging.info("\n" + str(sys.argv))
    logging.info("Test Data:" + FLAGS.test_dataset + " Training Data:" + FLAGS.dataset)
    util.print_num_of_total_parameters(model, output_to_logging=True)

    total_psnr = total_ssim = 0
    for i in range(FLAGS.tests):
        psnr, ssim = train(model, FLAGS, i)
        total_psnr += psnr
        total_ssim += ssim
        logging.info("\nTrial(%d) %s" % (i, util.get_now_date()))
        model.print_steps_completed(output_to_logging=True)
        logging.info("PSNR:%f, SSIM:%f\n" % (psnr, ssim))

    if FLAGS.tests > 1:
        logging.info("\n=== Final Average [%s] PSNR:%f, SSIM:%f ===" % (
            FLAGS.test_dataset, total_psnr / FLAGS.tests, total_ssim / FLAGS.tests))
    model.copy_log_to_archive("archive")


if __name__ == '__main__':
    args.run(main)
