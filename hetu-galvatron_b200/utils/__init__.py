from .config_utils import (array2str, config2strategy, read_json_config, str2array, strategy2config,
                           write_json_config)
from .strategy_utils import form_strategy, strategy_str2list
