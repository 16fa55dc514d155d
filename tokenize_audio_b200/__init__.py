"""B200-native Mimi audio tokenizer hot path (waveform -> codes), see DESIGN.md."""
