"""Only the names the reference imports (babyai/evaluate.py:3, scripts/train_rl.py:25).
RGBImgPartialObsWrapper restated from gym_minigrid 1.0.x (`wrappers.py`): the agent's 7x7 view rendered at tile_size
pixels per cell -- the 'pixel' architectures' input (scripts/train_rl.py:54-58, babyai/evaluate.py:91-92)."""
import gym
from gym import spaces


class RGBImgPartialObsWrapper(gym.ObservationWrapper):
    def __init__(self, env, tile_size=8):
        super().__init__(env)
        self.tile_size = tile_size
        obs_shape = env.observation_space.spaces['image'].shape
        self.observation_space.spaces['image'] = spaces.Box(
            low=0, high=255, shape=(obs_shape[0] * tile_size, obs_shape[1] * tile_size, 3), dtype='uint8')

    def observation(self, obs):
        env = self.unwrapped
        rgb_img_partial = env.get_obs_render(obs['image'], tile_size=self.tile_size)
        return {'mission': obs['mission'], 'image': rgb_img_partial}
