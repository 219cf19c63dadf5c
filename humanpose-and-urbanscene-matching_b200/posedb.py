"""Packed pose database: OpenPose JSON -> one flat file of [N][18][2] float32, streamed from disk to HBM.

The reference keeps one `*_keypoints.json` per image and parses it in Python for every comparison
(common.parse_JSON_multi_person, common.py:215-239; the retrieval loops of dataset/Multipose_dataset_actions.py:38-49).
Here the JSON is parsed once into a packed file; `PoseDB.search` memory-maps it and streams it through pinned
staging buffers on a copy stream while the previous chunk is being scored, keeping a running top-k on the device.

File format (little endian): magic b"HPUSMPDB", uint32 version = 1, uint32 joints = 18, uint64 n_poses,
uint64 n_files, then n_poses x 18 x 2 float32, then n_poses x uint32 file index, n_poses x uint32 person index, then the
file names as one UTF-8 blob of newline-separated names.
"""
import json
import os
import struct

import numpy as np
import torch

from . import engine

MAGIC = b"HPUSMPDB"
OPENPOSE_CERTAINTY = 0.2  # thresholds.OPENPOSE_ZEKERHEID (thresholds.py:3)


def parse_openpose_json(path, certainty=OPENPOSE_CERTAINTY):
    """List of (18,2) float64 arrays, one per person: joints whose confidence is not above `certainty` become (0,0)
    (common.parse_JSON_multi_person, common.py:215-239).  People without 18 x 3 values are skipped."""
    with open(path) as f:
        data = json.load(f)
    out = []
    for person in data.get("people", []):
        kp = person.get("pose_keypoints", person.get("pose_keypoints_2d"))
        if kp is None or len(kp) != 54:
            continue
        a = np.asarray(kp, dtype=np.float64).reshape(18, 3)
        pose = np.where(a[:, 2:3] > certainty, a[:, :2], 0.0)
        out.append(pose)
    return out


def pack(json_paths, out_path, certainty=OPENPOSE_CERTAINTY):
    """Parses the JSON files once and writes the packed database; returns the number of poses."""
    poses, file_idx, person_idx, names = [], [], [], []
    for fi, p in enumerate(json_paths):
        names.append(os.path.basename(p))
        for pi, pose in enumerate(parse_openpose_json(p, certainty)):
            poses.append(pose.astype(np.float32))
            file_idx.append(fi)
            person_idx.append(pi)
    arr = np.stack(poses) if poses else np.zeros((0, 18, 2), np.float32)
    blob = "\n".join(names).encode("utf-8")
    with open(out_path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<IIQQ", 1, 18, len(arr), len(names)))
        f.write(np.ascontiguousarray(arr, "<f4").tobytes())
        f.write(np.asarray(file_idx, "<u4").tobytes())
        f.write(np.asarray(person_idx, "<u4").tobytes())
        f.write(blob)
    return len(arr)


def pack_array(poses, out_path):
    """Writes an in-memory [N,18,2] array as a packed database (synthetic databases, tests)."""
    arr = np.ascontiguousarray(poses, "<f4").reshape(-1, 18, 2)
    with open(out_path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<IIQQ", 1, 18, len(arr), 0))
        f.write(arr.tobytes())
        f.write(np.zeros(len(arr), "<u4").tobytes())
        f.write(np.arange(len(arr), dtype="<u4").tobytes())
    return len(arr)


class PoseDB(object):
    HEADER = 8 + 4 + 4 + 8 + 8

    def __init__(self, path):
        with open(path, "rb") as f:
            head = f.read(self.HEADER)
        if head[:8] != MAGIC:
            raise ValueError("%s is not a packed pose database" % path)
        version, joints, n, nfiles = struct.unpack("<IIQQ", head[8:])
        if version != 1 or joints != 18:
            raise ValueError("unsupported pose database version %d / %d joints" % (version, joints))
        self.path, self.n, self.n_files = path, int(n), int(nfiles)
        self.poses = np.memmap(path, dtype="<f4", mode="r", offset=self.HEADER, shape=(self.n, 18, 2))
        meta = self.HEADER + self.n * 144
        self.file_index = np.memmap(path, dtype="<u4", mode="r", offset=meta, shape=(self.n,))
        self.person_index = np.memmap(path, dtype="<u4", mode="r", offset=meta + 4 * self.n, shape=(self.n,))
        with open(path, "rb") as f:
            f.seek(meta + 8 * self.n)
            blob = f.read().decode("utf-8")
        self.names = blob.split("\n") if blob else []

    def __len__(self):
        return self.n

    def search(self, query, k=16, device=None, chunk=1 << 21, thresholds=engine.DEFAULT_POSE_THRESHOLDS, query_is_model=True):
        """k best matching poses of the whole database: (scores [k] f32, indices [k] i64, match count) on the host.
        Disk -> pinned buffer -> HBM copies of chunk i+1 overlap the scoring of chunk i (two buffers, one copy stream)."""
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        q = torch.as_tensor(np.asarray(query, np.float32).reshape(18, 2)).to(device)
        copy_stream = torch.cuda.Stream(device)
        main = torch.cuda.current_stream(device)
        chunk = max(1, min(int(chunk), max(self.n, 1)))
        pinned = [torch.empty((chunk, 18, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
        dbuf = [torch.empty((chunk, 18, 2), dtype=torch.float32, device=device) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        host_free = [torch.cuda.Event() for _ in range(2)]
        best_s = torch.full((k,), float("inf"), dtype=torch.float32, device=device)
        best_i = torch.full((k,), -1, dtype=torch.int64, device=device)
        total = torch.zeros((1,), dtype=torch.int64, device=device)
        starts = list(range(0, self.n, chunk))

        def stage(ci):
            b = ci & 1
            lo = starts[ci]
            m = min(chunk, self.n - lo)
            if ci >= 2:
                host_free[b].synchronize()            # the H2D copy that last used this pinned buffer has finished
            pinned[b][:m].numpy()[...] = self.poses[lo:lo + m]  # disk (page cache) -> pinned
            with torch.cuda.stream(copy_stream):
                if ci >= 2:
                    copy_stream.wait_event(freed[b])  # the kernels that read this device buffer have finished
                dbuf[b][:m].copy_(pinned[b][:m], non_blocking=True)
                host_free[b].record(copy_stream)
                ready[b].record(copy_stream)
            return m

        sizes = {}
        if starts:
            sizes[0] = stage(0)
        for ci, lo in enumerate(starts):
            b = ci & 1
            if ci + 1 < len(starts):
                sizes[ci + 1] = stage(ci + 1)
            main.wait_event(ready[b])
            m = sizes[ci]
            match, score = engine.pose_match_batch(q, dbuf[b][:m], query_is_model, thresholds)
            ts, ti, cnt = engine.pose_topk(match, score, k, lo)
            freed[b].record(main)
            s = torch.cat([best_s, ts])
            i = torch.cat([best_i, ti])
            s = torch.where(i >= 0, s, torch.full_like(s, float("inf")))
            by_idx = torch.argsort(torch.where(i >= 0, i, torch.full_like(i, 1 << 62)), stable=True)
            order = by_idx[torch.argsort(s[by_idx], stable=True)][:k]
            best_s, best_i = s[order], i[order]
            total += cnt.to(torch.int64)
        torch.cuda.synchronize(device)
        return best_s.cpu().numpy(), best_i.cpu().numpy(), int(total.item())
