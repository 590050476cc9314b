# -*- coding: utf-8 -*-
"""Command line of the B200 NISQA engine - same flags and modes as the reference
``run_predict.py`` (reference run_predict.py:8-43): predict_file / predict_dir / predict_csv.

    python run_predict.py --mode predict_file --pretrained_model weights/nisqa.tar --deg a.wav
    python run_predict.py --mode predict_dir  --pretrained_model weights/nisqa.tar --data_dir d --bs 64
    torchrun --nproc-per-node 8 run_predict.py --mode predict_csv ... (one rank per GPU)
"""
import argparse

from nisqa_b200.NISQA_model import nisqaModel


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--mode", required=True, type=str, help="either predict_file, predict_dir, or predict_csv")
    p.add_argument("--pretrained_model", required=True, type=str, help="file name of pretrained model (must be in current working folder)")
    p.add_argument("--deg", type=str, help="path to speech file")
    p.add_argument("--data_dir", type=str, help="folder with speech files")
    p.add_argument("--output_dir", type=str, help="folder to ouput results.csv")
    p.add_argument("--csv_file", type=str, help="file name of csv (must be in current working folder)")
    p.add_argument("--csv_deg", type=str, help="column in csv with files name/path")
    p.add_argument("--num_workers", type=int, default=0, help="number of wav-decode worker threads")
    p.add_argument("--bs", type=int, default=1, help="batch size for predicting")
    p.add_argument("--ms_channel", type=int, help="audio channel in case of stereo file")
    args = vars(p.parse_args(argv))

    mode = args["mode"]
    if mode == "predict_file":
        if args["deg"] is None:
            raise ValueError("--deg argument with path to input file needed")
    elif mode == "predict_dir":
        if args["data_dir"] is None:
            raise ValueError("--data_dir argument with folder with input files needed")
    elif mode == "predict_csv":
        if args["csv_file"] is None:
            raise ValueError("--csv_file argument with csv file name needed")
        if args["csv_deg"] is None:
            raise ValueError("--csv_deg argument with csv column name of the filenames needed")
        if args["data_dir"] is None:
            args["data_dir"] = ""
    else:
        raise NotImplementedError("--mode given not available")
    args["tr_bs_val"] = args["bs"]
    args["tr_num_workers"] = args["num_workers"]
    return args


if __name__ == "__main__":
    import os
    from nisqa_b200 import dist as nb_dist
    # the CLI process only feeds the GPU: keep it (and its pinned batch buffers) on the GPU's NUMA node
    nb_dist.bind_to_gpu_numa(int(os.environ.get("LOCAL_RANK", "0")))
    nisqa = nisqaModel(parse_args())
    nisqa.predict()
