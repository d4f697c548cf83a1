"""CLI with the reference's flags (main_yelp.py / main_news.py of zyang1580/SML), running the B200 path.

    python main_yelp.py --data_path /data/ --data_name yelp --pre_model pre.pt --MF_epochs=1 --TR_epochs=1 --multi_num=10
    python main_yelp.py --profile news ...        # Adressa settings of main_news.py (63 periods, train from 21, test from 48)
"""
import argparse

import numpy as np
import torch

# flag -> (type, yelp default, news default, help); the names and defaults are the reference's
# (main_yelp.py:10-120, main_news.py:8-115)
FLAGS = [
    ("data_name", str, "yelp", "news", "dataset name"), ("data_path", str, "/home/sml/dataset/", "/home/sml/dataset/", "dataset path"),
    ("multi_num", int, 10, 7, "outer loop count (stop condition of SML)"),
    ("MF_lr", float, 0.01, 0.01, "learning rate of the MF step"), ("MF_epochs", int, 1, 2, "epochs of the MF step"),
    ("l2", float, 1e-6, 1e-6, "lambda_1"), ("MF_batch_size", int, 1024, 1024, "MF batch"), ("laten", int, 64, 64, "embedding dim"),
    ("pre_model", str, "/home/sml/save_model/sml/yelp/BCE_init.pkl", "/home/sml/save_model/sml/news/BCE_init.pkl",
     "pre-trained MF (pickled module or state_dict)"), ("MF_sample", str, "all", "all", "all | alone"),
    ("Load_W_hat", bool, False, False, ""), ("clip_grad", bool, False, False, ""), ("need_adaptive", bool, False, False, ""),
    ("maxnorm_grad", float, 3.0, 3.0, ""),
    ("TR_lr", float, 0.001, 0.001, "transfer learning rate"), ("TR_l2", float, 0.0001, 0.0001, "lambda_2"),
    ("TR_epochs", int, 1, 2, "epochs of the transfer step"), ("TR_batch_size", int, 256, 256, "transfer batch"),
    ("TR_sample_type", str, "alone", "alone", "all | alone"), ("TR_with_MF_bias", bool, False, False, ""), ("TR_stop_", bool, False, False, ""),
    ("transfer_type", str, "conv_com", "conv_com", "conv_com | conv"),
    ("seed", int, 2000, 2000, ""), ("numworkers", int, 4, 4, "accepted, unused: batches are built on the host at once"),
    ("cuda", int, 0, 0, "GPU index"), ("topK", int, 20, 20, ""), ("pass_num", int, 1, 1, ""), ("norm", bool, False, False, ""),
    ("Lambda_lr", float, 0.01, 0.01, ""), ("min_l2", float, 0.0001, 0.0001, ""), ("set_t_as_tt", bool, False, False, ""),
    ("tqdm", bool, False, False, ""), ("need_writer", bool, False, False, ""), ("test_in_TR_Train", bool, False, False, ""),
]
STREAMS = {"yelp": dict(n_files=40, train_from=10, test_from=30), "news": dict(n_files=63, train_from=21, test_from=48)}


def get_parse(profile="yelp"):
    parser = argparse.ArgumentParser(description="MF and TR(transfer) parameters in SML (B200 path).")
    parser.add_argument("--profile", default=profile, choices=sorted(STREAMS), help="defaults + period lists of main_yelp.py or main_news.py")
    parser.add_argument("--device_sampler", action="store_true", help="throughput runs: sample batches on the GPU (Philox)")
    col = 2 if profile == "yelp" else 3
    for f in FLAGS:
        # the reference declares its booleans with type=bool (any non-empty string is True); kept
        parser.add_argument("--" + f[0], type=f[1], default=f[col], help=f[4])
    return parser


def main(profile="yelp"):
    pre = argparse.ArgumentParser(add_help=False)
    pre.add_argument("--profile", default=profile)
    profile = pre.parse_known_args()[0].profile
    args = get_parse(profile).parse_args()
    torch.cuda.set_device(args.cuda)
    torch.manual_seed(args.seed); torch.cuda.manual_seed(args.seed + 1); np.random.seed(args.seed + 2)      # main_yelp.py:137-140
    from sml_b200.data import dataset2
    from sml_b200.model import transfer
    st = STREAMS[profile]
    file_list = [str(i) for i in range(st["n_files"])]
    test_list = [str(j) for j in range(st["test_from"], st["n_files"])]
    ds = dataset2.transfer_data(args, path=args.data_path, datasetname=args.data_name, file_path_list=file_list, test_list=test_list,
                                validation_list=None, online_train_time=st["train_from"], online_test_time=st["test_from"])
    meta = transfer.meta_train(args, ds, ds.user_number, ds.item_number, args.laten, device_sampler=args.device_sampler)
    meta.run(args)
    return meta


if __name__ == "__main__":
    main("yelp")
