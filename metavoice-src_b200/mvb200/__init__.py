"""mvb200: B200-native engine behind MetaVoice-1B's TTS.synthesise() hot path."""
