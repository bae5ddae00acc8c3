from .utils import (split_channel_name, unique_list, convert_length_msec_to_samples,
                    get_window_indices)
