"""Packed image database: descriptors + keypoints of many images in one flat file, streamed from disk to HBM.

The reference re-detects and re-describes both images for every pair it compares (dataset/dataset_urbanscene.py:46-84
loops urban_scene.match_scene_* over image pairs; urban_scene.py:31-32 calls detector.detectAndCompute each time).  Here
every database image is described ONCE (`pack_images`, with the detector parameters of features.init_feature,
features.py:10-43) into a packed file; `ImageDB.retrieve` memory-maps it and streams runs of whole images through pinned
staging buffers on a copy stream while the previous run goes through retrieval.scene_retrieval (segmented Hamming kNN-2 +
ratio -> per-image match lists -> batched RANSAC both ways -> homography sanity), so a query image is matched against the
whole database without a host round trip per image.

File format (little endian): magic b"HPUSMIDB", uint32 version = 1, uint32 kind (0 = binary descriptors, uint8 rows),
uint32 width (bytes per row: 32 for ORB, 64 for BRISK / padded AKAZE), uint32 reserved, uint64 n_images, uint64 n_rows;
then (n_images + 1) x int64 row offsets, n_rows x width uint8 descriptors, n_rows x 2 float32 keypoint coordinates
(KeyPoint.pt), n_images x 2 uint32 image sizes (height, width), then the image names as one UTF-8 blob, newline separated.
"""
import os
import struct

import numpy as np
import torch

from . import retrieval

MAGIC = b"HPUSMIDB"
KIND_BINARY = 0


def _pad_width(desc):
    """Binary descriptors zero-padded to the kernel's row widths (32 or 64 bytes; AKAZE's MLDB rows are 61 bytes)."""
    desc = np.ascontiguousarray(desc, dtype=np.uint8)
    w = desc.shape[1]
    if w in (32, 64):
        return desc
    if w > 64:
        raise ValueError("binary descriptors wider than 64 bytes are not supported (got %d)" % w)
    out = np.zeros((desc.shape[0], 32 if w < 32 else 64), np.uint8)
    out[:, :w] = desc
    return out


def pack_arrays(descriptors, keypoints, out_path, names=None, sizes=None):
    """Writes per-image descriptor / keypoint arrays as a packed database; returns (n_images, n_rows).
    descriptors[i]: [n_i, width] uint8 (or None for an image without keypoints); keypoints[i]: [n_i, 2] float32."""
    nimg = len(descriptors)
    descs, kps, offs = [], [], [0]
    width = None
    for d, k in zip(descriptors, keypoints):
        if d is None or len(d) == 0:
            offs.append(offs[-1])
            continue
        d = _pad_width(d)
        if width is None:
            width = d.shape[1]
        if d.shape[1] != width:
            raise ValueError("all images must share one descriptor width (got %d and %d)" % (width, d.shape[1]))
        k = np.ascontiguousarray(k, dtype="<f4").reshape(-1, 2)
        if len(k) != len(d):
            raise ValueError("descriptor / keypoint counts differ: %d vs %d" % (len(d), len(k)))
        descs.append(d)
        kps.append(k)
        offs.append(offs[-1] + len(d))
    width = width or 32
    n_rows = offs[-1]
    names = list(names) if names is not None else ["image%d" % i for i in range(nimg)]
    sizes = np.asarray(sizes if sizes is not None else np.zeros((nimg, 2)), dtype="<u4").reshape(nimg, 2)
    with open(out_path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<IIIIQQ", 1, KIND_BINARY, width, 0, nimg, n_rows))
        f.write(np.asarray(offs, "<i8").tobytes())
        f.write((np.concatenate(descs) if descs else np.zeros((0, width), np.uint8)).tobytes())
        f.write((np.concatenate(kps) if kps else np.zeros((0, 2), "<f4")).tobytes())
        f.write(sizes.tobytes())
        f.write("\n".join(names).encode("utf-8"))
    return nimg, n_rows


def pack_images(image_paths, out_path, feature_name="orb"):
    """Detects and describes every image once with the detector of features.init_feature(feature_name) (binary
    descriptors: 'orb', 'akaze', 'brisk'; features.py:10-43) and writes the packed database."""
    import cv2
    from .dropin.urbanscene import features
    detector, matcher = features.init_feature(feature_name)
    if detector is None or getattr(matcher, "normType", None) != features.NORM_HAMMING:
        raise ValueError("pack_images: %r is not a binary-descriptor feature" % feature_name)
    descs, kps, names, sizes = [], [], [], []
    for p in image_paths:
        img = cv2.imread(p, 0)
        if img is None:
            raise IOError("cannot read image %s" % p)
        kp, desc = detector.detectAndCompute(img, None)
        descs.append(desc)
        kps.append(np.float32([k.pt for k in kp]).reshape(-1, 2))
        names.append(os.path.basename(p))
        sizes.append(img.shape[:2])
    return pack_arrays(descs, kps, out_path, names, sizes)


class ImageDB(object):
    HEADER = 8 + 4 * 4 + 8 + 8

    def __init__(self, path):
        with open(path, "rb") as f:
            head = f.read(self.HEADER)
        if head[:8] != MAGIC:
            raise ValueError("%s is not a packed image database" % path)
        version, kind, width, _, nimg, nrows = struct.unpack("<IIIIQQ", head[8:])
        if version != 1 or kind != KIND_BINARY or width not in (32, 64):
            raise ValueError("unsupported image database (version %d, kind %d, width %d)" % (version, kind, width))
        self.path, self.n_images, self.n_rows, self.width = path, int(nimg), int(nrows), int(width)
        off = self.HEADER
        self.seg_offsets = np.array(np.memmap(path, dtype="<i8", mode="r", offset=off, shape=(self.n_images + 1,)))
        off += 8 * (self.n_images + 1)
        self.descriptors = np.memmap(path, dtype=np.uint8, mode="r", offset=off, shape=(self.n_rows, self.width)) \
            if self.n_rows else np.zeros((0, self.width), np.uint8)
        off += self.n_rows * self.width
        self.keypoints = np.memmap(path, dtype="<f4", mode="r", offset=off, shape=(self.n_rows, 2)) \
            if self.n_rows else np.zeros((0, 2), np.float32)
        off += self.n_rows * 8
        self.sizes = np.array(np.memmap(path, dtype="<u4", mode="r", offset=off, shape=(self.n_images, 2))) \
            if self.n_images else np.zeros((0, 2), np.uint32)
        off += self.n_images * 8
        with open(path, "rb") as f:
            f.seek(off)
            blob = f.read().decode("utf-8")
        self.names = blob.split("\n") if blob else []

    def __len__(self):
        return self.n_images

    def image(self, i):
        """(descriptors [n_i, width] uint8, keypoints [n_i, 2] float32) of image i."""
        lo, hi = int(self.seg_offsets[i]), int(self.seg_offsets[i + 1])
        return self.descriptors[lo:hi], self.keypoints[lo:hi]

    def chunks(self, images_per_chunk):
        """Runs of whole images: (first image, last image + 1, first row, last row + 1)."""
        step = max(int(images_per_chunk), 1)
        for i0 in range(0, self.n_images, step):
            i1 = min(self.n_images, i0 + step)
            yield i0, i1, int(self.seg_offsets[i0]), int(self.seg_offsets[i1])

    def retrieve(self, query_desc, query_kp, device=None, images_per_chunk=1024, ratio=0.8, min_match=16, thresh=5.0,
                 nhyp=2000, confidence=0.995, seed=0, sampler="counter"):
        """One query image against every database image: features.match_and_draw + validate_homography per image
        (features.py:57-105,:177-211), streamed.  Returns host arrays, one entry per database image:
        n_matches, n_inliers (int32), valid (uint8), H, H2 ([n_images,3,3] float64, NaN where there is no model).
        Disk -> pinned -> HBM copies of run i+1 overlap the matching of run i (two buffers, one copy stream)."""
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        q = torch.as_tensor(_pad_width(np.asarray(query_desc))).to(device)
        if q.shape[1] != self.width:
            raise ValueError("query descriptors are %d bytes wide, the database %d" % (q.shape[1], self.width))
        qkp = torch.as_tensor(np.ascontiguousarray(query_kp, np.float32).reshape(-1, 2)).to(device)
        runs = list(self.chunks(images_per_chunk))
        cap = max([r1 - r0 for _, _, r0, r1 in runs] or [0])
        copy_stream = torch.cuda.Stream(device)
        main = torch.cuda.current_stream(device)
        pin_d = [torch.empty((max(cap, 1), self.width), dtype=torch.uint8).pin_memory() for _ in range(2)]
        pin_k = [torch.empty((max(cap, 1), 2), dtype=torch.float32).pin_memory() for _ in range(2)]
        dev_d = [torch.empty((max(cap, 1), self.width), dtype=torch.uint8, device=device) for _ in range(2)]
        dev_k = [torch.empty((max(cap, 1), 2), dtype=torch.float32, device=device) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        host_free = [torch.cuda.Event() for _ in range(2)]
        n_matches = torch.zeros((self.n_images,), dtype=torch.int32, device=device)
        n_inliers = torch.zeros((self.n_images,), dtype=torch.int32, device=device)
        valid = torch.zeros((self.n_images,), dtype=torch.uint8, device=device)
        H = torch.full((self.n_images, 3, 3), float("nan"), dtype=torch.float64, device=device)
        H2 = torch.full((self.n_images, 3, 3), float("nan"), dtype=torch.float64, device=device)

        def stage(ci):
            b = ci & 1
            _, _, r0, r1 = runs[ci]
            m = r1 - r0
            if ci >= 2:
                host_free[b].synchronize()            # the H2D copies that last used these pinned buffers have finished
            if m:
                pin_d[b][:m].numpy()[...] = self.descriptors[r0:r1]  # disk (page cache) -> pinned
                pin_k[b][:m].numpy()[...] = self.keypoints[r0:r1]
            with torch.cuda.stream(copy_stream):
                if ci >= 2:
                    copy_stream.wait_event(freed[b])  # the kernels that read these device buffers have finished
                if m:
                    dev_d[b][:m].copy_(pin_d[b][:m], non_blocking=True)
                    dev_k[b][:m].copy_(pin_k[b][:m], non_blocking=True)
                host_free[b].record(copy_stream)
                ready[b].record(copy_stream)

        if runs:
            stage(0)
        for ci, (i0, i1, r0, r1) in enumerate(runs):
            b = ci & 1
            if ci + 1 < len(runs):
                stage(ci + 1)
            main.wait_event(ready[b])
            m = r1 - r0
            if m:
                local = torch.as_tensor(self.seg_offsets[i0:i1 + 1] - r0).to(device)
                out = retrieval.scene_retrieval(q, qkp, dev_d[b][:m], dev_k[b][:m], local, ratio, min_match, thresh, nhyp,
                                                confidence, seed, sampler=sampler, image_index0=i0)
                n_matches[i0:i1], n_inliers[i0:i1], valid[i0:i1] = out["n_matches"], out["n_inliers"], out["valid"]
                H[i0:i1], H2[i0:i1] = out["H"], out["H2"]
            freed[b].record(main)
        torch.cuda.synchronize(device)
        return {"n_matches": n_matches.cpu().numpy(), "n_inliers": n_inliers.cpu().numpy(), "valid": valid.cpu().numpy(),
                "H": H.cpu().numpy(), "H2": H2.cpu().numpy()}
