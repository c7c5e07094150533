"""Same path and name as the reference's entry point; forwards to the B200-native implementation.

    python similar_face_filtering/filter_faces_using_reference.py --ud U --rd R --td T [-m W] [-b 32] [-r 32]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from face_detection_and_recognition_b200.filter_faces_using_reference import (  # noqa: E402,F401
    _fix_path_for_globbing, get_class_name_list, get_parsed_args, get_ref_mean_vec_and_thres_from_imgs, main,
    read_and_preprocess_img)

if __name__ == "__main__":
    main()
